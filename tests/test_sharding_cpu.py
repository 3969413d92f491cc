"""World-size-2 gloo test of the multi-GPU host logic (frame-range sharding + result reassembly)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from humanoid_real_time_retarget_b200.sharding import all_gather_frames, shard_range
    lo, hi = shard_range(n_frames, rank, world)
    full = torch.arange(n_frames * 30, dtype=torch.float32).reshape(n_frames, 30)
    local = full[lo:hi] * 2.0                      # stand-in for the per-rank kernel output
    out = all_gather_frames(local, n_frames)
    ok = torch.equal(out, full * 2.0)
    ret[rank] = (lo, hi, bool(ok))
    dist.destroy_process_group()


def test_shard_ranges_cover_the_clip_exactly():
    from humanoid_real_time_retarget_b200.sharding import shard_range
    for n in (0, 1, 15, 16, 17, 1000, 1 << 20, (1 << 24) + 5):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all((lo % 16 == 0) for lo, _ in spans if lo < n)


def test_reassembly_transport_by_rank_count():
    from humanoid_real_time_retarget_b200.sharding import reassembly_transport as rt
    assert [rt("auto", w, True) for w in (1, 2, 3, 4, 5, 8)] == ["unicast", "multicast", "multicast", "multicast", "packed", "packed"]
    assert [rt("auto", w, False) for w in (2, 8)] == ["unicast", "unicast"]
    assert rt("packed", 2, True) == "packed" and rt("multicast", 8, True) == "multicast" and rt("unicast", 8, True) == "unicast"
    with pytest.raises(RuntimeError):
        rt("packed", 8, False)
    with pytest.raises(ValueError):
        rt("nccl", 8, True)


def test_all_gather_reassembles_ragged_shards_gloo():
    world, n_frames = 2, 1000 + 7                  # ragged: rank 1 holds fewer frames
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_frames, ret), nprocs=world, join=True)
    assert ret[0][2] and ret[1][2]
    assert ret[0][0] == 0 and ret[0][1] == ret[1][0] and ret[1][1] == n_frames


class _FakeEngine:
    """Stands in for the CUDA engine on the CPU: dof_pos[f] = 2 * (first 30 words of frame f)."""

    def retarget_body_quat(self, raw, flags=0, ik_iters=0, damping=0.0, rot_weight=0.0, out=None):
        out[1].copy_(raw.reshape(raw.shape[0], -1)[:, :30] * 2.0)
        return out


def _worker_cyclic(rank, world, port, n_frames, n_blocks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from humanoid_real_time_retarget_b200.sharding import block_cyclic_ranges, retarget_clip_overlapped
    full = torch.arange(n_frames * 84, dtype=torch.float32).reshape(n_frames, 21, 4)
    blk, spans = block_cyclic_ranges(n_frames, rank, world, n_blocks)
    mine = torch.zeros(n_blocks, blk, 21, 4)
    for c, (lo, hi) in enumerate(spans):
        mine[c, : hi - lo] = full[lo:hi]
    dof, _ = retarget_clip_overlapped(_FakeEngine(), mine, n_frames, flags=0, n_blocks=n_blocks)
    ret[rank] = bool(torch.equal(dof, full.reshape(n_frames, -1)[:, :30] * 2.0))
    dist.destroy_process_group()


def test_block_cyclic_overlapped_gather_lands_in_frame_order_gloo():
    from humanoid_real_time_retarget_b200.sharding import block_cyclic_ranges
    for n, w, nb in ((1000, 2, 4), (1 << 20, 8, 4), (17, 4, 2)):
        blk, _ = block_cyclic_ranges(n, 0, w, nb)
        covered = sorted(s for r in range(w) for s in block_cyclic_ranges(n, r, w, nb)[1] if s[1] > s[0])
        assert covered[0][0] == 0 and covered[-1][1] == n and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        assert blk % 16 == 0
    world, n_frames, n_blocks = 2, 1000 + 7, 4
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_cyclic, args=(world, _free_port(), n_frames, n_blocks, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]


class _ScipyEngine:
    """CPU stand-in with the reference's arithmetic (np.gradient + gaussian_filter1d, skeleton3d.py:1126-1135)."""

    def motion_velocity(self, gt, dt, gaussian=True):
        import numpy as np
        from scipy.ndimage import gaussian_filter1d
        v = np.gradient(gt.numpy(), axis=-3) / dt
        return torch.from_numpy(gaussian_filter1d(v, 2, axis=-3, mode="nearest") if gaussian else v)

    def motion_angular_velocity(self, gq, dt, gaussian=True):
        return self.motion_velocity(gq[..., :3], dt, gaussian)            # any frame-local map + the same stencil


def _worker_halo(rank, world, port, n_frames, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from humanoid_real_time_retarget_b200.sharding import exchange_halo, motion_velocities_sharded, shard_range
    g = torch.Generator().manual_seed(5)
    gt = torch.cumsum(torch.randn(n_frames, 4, 3, generator=g), dim=0)
    gq = torch.randn(n_frames, 4, 4, generator=g)
    lo, hi = shard_range(n_frames, rank, world)
    padded, lead = exchange_halo(gt[lo:hi], 9)
    ok_halo = torch.equal(padded, gt[max(0, lo - 9):min(n_frames, hi + 9)]) and lead == (9 if rank else 0)
    eng = _ScipyEngine()
    vel, ang = motion_velocities_sharded(eng, gt[lo:hi], gq[lo:hi], 1 / 30)
    ok_vel = torch.equal(vel, eng.motion_velocity(gt, 1 / 30)[lo:hi]) and torch.equal(ang, eng.motion_angular_velocity(gq, 1 / 30)[lo:hi])
    try:
        exchange_halo(gt[:4], 9)
        short = False
    except ValueError:
        short = True
    ret[rank] = (bool(ok_halo), bool(ok_vel), short)
    dist.destroy_process_group()


def test_velocity_halo_exchange_matches_whole_clip_gloo():
    """a 9-frame halo from each neighbour makes the sharded velocities equal the whole-clip ones."""
    world, n_frames = 2, 203
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_halo, args=(world, _free_port(), n_frames, ret), nprocs=world, join=True)
    assert all(all(ret[r]) for r in range(world)), dict(ret)


def test_gpu_local_host_memory_never_raises_without_a_gpu():
    """The NUMA helper is best effort: without NVML / a GPU it reports why it did nothing and leaves the affinity alone."""
    from humanoid_real_time_retarget_b200.sharding import gpu_local_host_memory
    before = os.sched_getaffinity(0)
    with gpu_local_host_memory(0) as h:
        assert isinstance(h.info, dict) and ("skipped" in h.info or "cpus" in h.info)
    assert os.sched_getaffinity(0) == before
