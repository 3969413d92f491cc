"""Single-GPU cost of the in-kernel unpacking (for ncu and quick timing): the GATHER instantiation runs with 7 fake peers whose
flags count as raised (HRT_GATHER_DEBUG bit 32) and whose staging rows are a plain local buffer; nothing is sent
(bits 1 | 16).  Results are meaningless, the instruction stream of fetch + expand is the real one.
    HRT_GATHER_DEBUG=49 python tools/gather_unpack_probe.py [n_rank] [frames per rank]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HRT_GATHER_DEBUG", "49")
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402
from oracle import retarget_oracle as oc  # noqa: E402  (input synthesis only)

n_rank = int(sys.argv[1]) if len(sys.argv) > 1 else 8
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 21
eng = hrt.Engine(0).set_standard_trees()
sk = oc.load_skeletons()
raw = oc.synth_clip_3q(1 << 16, seed=5, sk=sk).cuda().repeat(B >> 16, 1, 1).contiguous()
flags = hrt.BQ_CLAMP | hrt.BQ_IK
n_total = n_rank * B
shard_n = [B] * n_rank
shard_lo = [r * B for r in range(n_rank)]
_, _, total, _ = eng.reassembly_layout(n_total, shard_n)
symm = torch.zeros(total // 4, device="cuda")
full = torch.zeros(n_total, 30, device="cuda")
lp = torch.empty(B, 31, 3, device="cuda")
dof = torch.empty(B, 30, device="cuda")


def timed(fn, k=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(k):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / k


ep = [0]


def gather_step():
    ep[0] += 1
    eng.retarget_body_quat_reassemble(raw, full, n_total, 0, shard_lo, shard_n, symm.data_ptr(), symm.data_ptr(), ep[0], flags=flags, link_pos=lp)


plain = timed(lambda: eng.retarget_body_quat(raw, flags=flags, out=(None, dof, lp)))
g = timed(gather_step)
print(f"n_rank={n_rank} frames/rank={B} HRT_GATHER_DEBUG={os.environ['HRT_GATHER_DEBUG']}: plain {plain:.4f} ms, with unpack of {n_rank - 1} peers {g:.4f} ms "
      f"(+{100 * (g / plain - 1):.1f} %)")
