"""N-rank check of the fused reassembly (sharding.PeerReassembly): run under torchrun on N GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_gather_check.py

Every rank retargets its shard of a clip whose frames are generated from one seed (so every rank can compute the whole
answer locally), the kernels store their dof spans into every rank's buffer over NVLink, and every rank compares the
reassembled clip bit for bit with its own single-GPU result.  Ragged clip lengths included."""
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
    ok = True
    for transport in ("auto", "multicast", "unicast"):
        for n in (1 << 18, 100_003, 16 * world * 3 + 5, 7):
            raw = oc.synth_clip_3q(n, seed=5, sk=sk).cuda()
            lo, hi = shard_range(n, rank, world)
            pr = PeerReassembly(eng, n, transport=transport)
            for _ in range(3):
                full = pr.step(raw[lo:hi], flags)
            torch.cuda.synchronize()
            dist.barrier()
            _, want, _ = eng.retarget_body_quat(raw, flags=flags, want_local_q=False, want_link_pos=False)
            same = bool(torch.equal(full, want))
            print(f"rank {rank}/{world} n={n} transport={pr.transport} ({pr.transport_error}): shard [{lo},{hi}) reassembled clip "
                  f"bit-equal to the local result: {same}", flush=True)
            ok &= same
            pr.close()
    t = torch.tensor([float(ok)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if t.item() != 1.0:
        sys.exit(1)


if __name__ == "__main__":
    main()
