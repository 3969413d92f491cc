"""Where does the in-kernel reassembly spend its time?  Run under torchrun on 2+ GPUs, once per HRT_GATHER_DEBUG mask:
    for m in 0 7 31; do HRT_GATHER_DEBUG=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29515 tools/gather_probe.py; done
Masks (diagnostics only, results are wrong with any of them): 1 no packed data stores, 2 no unpack stores, 4 no unpack loads,
8 relaxed instead of release flag store."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    g.build()
    import humanoid_real_time_retarget_b200 as hrt
    from humanoid_real_time_retarget_b200.sharding import PeerReassembly, shard_range
    from oracle import retarget_oracle as oc
    sk = oc.load_skeletons()
    eng = hrt.Engine(local).set_standard_trees()
    flags = hrt.BQ_CLAMP | hrt.BQ_IK
    n = int(os.environ.get("PROBE_FRAMES", 1 << 22))
    lo, hi = shard_range(n, rank, world)
    raw = oc.synth_clip_3q(1 << 16, seed=5 + rank, sk=sk).cuda().repeat((hi - lo + 65535) // 65536, 1, 1)[: hi - lo].contiguous()
    lp = torch.empty(hi - lo, 31, 3, device="cuda")
    dof = torch.empty(hi - lo, 30, device="cuda")
    res = {}
    for name in ("plain", "packed"):
        pr = PeerReassembly(eng, n, transport="packed") if name == "packed" else None
        step = (lambda: pr.step(raw, flags, link_pos=lp)) if pr else (lambda: eng.retarget_body_quat(raw, flags=flags, out=(None, dof, lp)))
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            step()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 10], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = round(float(t.item()), 4)
        if pr:
            pr.close()
    if rank == 0:
        print(f"HRT_GATHER_DEBUG={os.environ.get('HRT_GATHER_DEBUG', '0')} world={world} frames={n}: {res}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
