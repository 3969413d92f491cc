"""Where does position-path disagreement with the reference come from?  (CPU only, dev container or GPU box.)

    python tools/parity_study.py            ->  profiles/parity_study_cpu_r02.json

The oracle (oracle/retarget_oracle.py) restates the reference op for op and reproduces its per-frame outputs on the golden
frames (tests/golden/*.npz, written by the UNMODIFIED reference).  This script swaps ONE ingredient of the oracle at a time
and measures how far the result moves from the reference's goldens:

  oracle          the oracle as is: torch fp32 ops, Kabsch through torch.linalg.svd = MKL sgesdd in fp32
  svd_fp64        Kabsch rotation from a float64 SVD, rounded once to fp32 (what an exact 3x3 SVD returns; the CUDA
                  kernels' fp64 solve is of this kind).  Everything else unchanged.  The distance of THIS variant from the
                  reference is the part of the disagreement that no implementation without MKL's sgesdd can remove.
  libm_cr         acos / atan2 / sin / cos / sqrt evaluated in float64 and rounded once (correctly rounded fp32 results);
                  SVD unchanged (MKL).  Measures how much the choice of libm matters.
  svd_fp64+libm   both.

For every variant: fraction of frames whose worst |d dof| <= 1e-5 rad, p50 / p99 / max, the FK link-position error of the
two angle sets (FK is well conditioned where the raw angles are not) and the geodesic error of the local rotations.
"""
import contextlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import retarget_oracle as oc  # noqa: E402
from oracle import parity_metrics as pm  # noqa: E402

T = torch.from_numpy
GOLD = os.path.join(ROOT, "tests", "golden")


kabsch_fp64 = oc.cal_joint_quat_exact_svd


@contextlib.contextmanager
def patched(svd64=False, libm=False):
    saved = {}
    if svd64:
        saved["cjq"] = oc.cal_joint_quat
        oc.cal_joint_quat = kabsch_fp64
    names = ["acos", "atan2", "sin", "cos", "sqrt"]
    if libm:
        for n in names:
            saved[n] = getattr(torch, n)
        cr = lambda f: (lambda *a: f(*[x.double() for x in a]).float())
        for n in names:
            setattr(torch, n, cr(saved[n]))
        saved["t_sin"], saved["t_cos"] = torch.Tensor.sin, torch.Tensor.cos
        torch.Tensor.sin = lambda self: saved["t_sin"](self.double()).float() if self.dtype == torch.float32 else saved["t_sin"](self)
        torch.Tensor.cos = lambda self: saved["t_cos"](self.double()).float() if self.dtype == torch.float32 else saved["t_cos"](self)
    try:
        yield
    finally:
        if svd64:
            oc.cal_joint_quat = saved["cjq"]
        if libm:
            for n in names:
                setattr(torch, n, saved[n])
            torch.Tensor.sin, torch.Tensor.cos = saved["t_sin"], saved["t_cos"]


def stats(dof, rl, dof_ref, rl_ref, sk):
    return pm.distance_stats(dof, dof_ref, rl, rl_ref, sk["hu_v5_zero_pose/parents"].tolist(), sk["hu_v5_zero_pose/offsets"])


def main():
    sk = oc.load_skeletons()
    out = {"what": __doc__.split("\n\n")[1].strip(), "lapack": "MKL 2024.2 sgesdd via torch.linalg.svd (closed source)", "sets": {}}
    g = np.load(os.path.join(GOLD, "full_body_pos.npz"))
    off59, zgt = T(sk["vtrdyn_full_zero_pose/offsets"]), T(sk["vtrdyn_full_zero_pose/global_translation"])
    gu = np.load(os.path.join(GOLD, "upper_body.npz"))
    off21 = T(sk["vtrdyn_zero_pose/offsets"])
    gf = np.load(os.path.join(GOLD, "full_body.npz"))
    variants = [("oracle", {}), ("svd_fp64", {"svd64": True}), ("libm_cr", {"libm": True}), ("svd_fp64+libm", {"svd64": True, "libm": True})]
    for name, kw in variants:
        with patched(**kw):
            rl, dof, _ = oc.retarget_full_body_pos(T(g["body_t"]), T(g["lhand_t"]), T(g["rhand_t"]), off59, zgt, True)
            out["sets"].setdefault("full_body_pos (a32, 256 golden frames)", {})[name] = stats(dof, rl, T(g["dof_pos"]), T(g["robot_local_q"]), sk)
            rl, dof = oc.retarget_upper_body(T(gu["global_t"]), off21)
            out["sets"].setdefault("upper_body (a29, 256 golden frames)", {})[name] = stats(dof, rl, T(gu["dof_pos"]), T(gu["robot_local_q"]), sk)
            rl, dof = oc.retarget_full_body(T(gf["body_q"]), T(gf["body_t"]), T(gf["lhand_t"]), T(gf["rhand_t"]), off59)
            out["sets"].setdefault("full_body (a31, 256 golden frames; no Kabsch on this path)", {})[name] = stats(dof, rl, T(gf["dof_pos"]), T(gf["robot_local_q"]), sk)
    # the reference against itself under 1-ulp input jitter, stored with the goldens (tools/make_golden.py)
    sd = g["self_delta"]
    out["reference_self_delta_1ulp_jitter (a32)"] = {"frac_le_1e-5": float((sd <= 1e-5).mean()), "p50": float(np.median(sd)),
                                                     "p99": float(np.quantile(sd, .99)), "max": float(sd.max())}
    # Kabsch primitive on its own: quaternion component error vs the reference's (MKL fp32) result
    p = np.load(os.path.join(GOLD, "primitives.npz"))
    for key, Z, M in (("kabsch3", "Z3", "M3"), ("kabsch5", "Z5", "M5")):
        q64 = kabsch_fp64(T(p[Z]), T(p[M]))
        e = (q64 - T(p[key])).abs().max(dim=-1).values.numpy()
        out.setdefault("kabsch_fp64_vs_reference_quat", {})[key] = {"p50": float(np.median(e)), "p99": float(np.quantile(e, .99)), "max": float(e.max())}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    path = os.path.join(ROOT, "profiles", "parity_study_cpu_r02.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
