"""Diagnostic (not a test): per-frame error of the position-path kernel against the oracle next to the
oracle's own sensitivity to 1-ulp input jitter (several probes).  Writes gpurun_out/diag_pos.npz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402

T = torch.from_numpy


def main(B=20000, probes=6):
    sk = oc.load_skeletons()
    g = torch.Generator().manual_seed(0)
    em = 0.4 * torch.randn(B, 59, 3, generator=g)
    lq = oc.exp_map_to_quat(em)
    root = torch.zeros(B, 3)
    root[:, 2] = 1.0
    _, gt = oc.cal_forward_kinematics(lq, root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                      T(sk["vtrdyn_full_zero_pose/offsets"]))
    full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, full2body].contiguous(), gt[:, 14:34].contiguous(), gt[:, 39:59].contiguous()
    off = T(sk["vtrdyn_full_zero_pose/offsets"])
    zgt = T(sk["vtrdyn_full_zero_pose/global_translation"])
    _, dof_o, _ = oc.retarget_full_body_pos(body, lh, rh, off, zgt, True)
    deltas = []
    for p in range(probes):
        gj = torch.Generator().manual_seed(1 + p)
        jit = lambda x: torch.nextafter(x, x + torch.sign(torch.randn(x.shape, generator=gj)))
        _, dof_j, _ = oc.retarget_full_body_pos(jit(body), jit(lh), jit(rh), off, zgt, True)
        deltas.append((dof_j - dof_o).abs().numpy())
    eng = hrt.default_engine(0)
    _, dof_k, _ = eng.retarget_full_body_pos(body, lh, rh)
    err = (dof_k.cpu() - dof_o).abs().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "diag_pos.npz"), err=err, deltas=np.stack(deltas),
                        dof_o=dof_o.numpy(), dof_k=dof_k.cpu().numpy())
    e = err.max(-1)
    d = np.stack(deltas).max(-1)
    for n in range(1, probes + 1):
        cond = d[:n].max(0) < 2e-6
        print(f"probes={n}: coverage {cond.mean():.3f} kernel within 1e-5 on it {(e[cond] <= 1e-5).mean():.4f}")


if __name__ == "__main__":
    main()
