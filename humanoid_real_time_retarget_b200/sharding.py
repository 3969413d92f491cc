"""Frame-range sharding of a clip across the GPUs of one box.

Frames are independent, so the data path needs no collective: rank r of W processes frames
[r*ceil(B/W), min(B, (r+1)*ceil(B/W))).  The only exchange the path ever needs is the optional
reassembly of the results on every rank: one all-gather over NCCL / NVLink (gloo on CPU in the tests)."""
import torch
import torch.distributed as dist


def shard_range(n_frames, rank, world):
    """Contiguous, balanced to within one 16-frame warp group (the kernels' staging granularity)."""
    per = -(-n_frames // world)
    per = -(-per // 16) * 16
    lo = min(n_frames, rank * per)
    hi = min(n_frames, lo + per)
    return lo, hi


def all_gather_frames(local, n_frames, group=None):
    """Reassemble per-rank results (leading axis = this rank's frames) into the full clip on every rank.
    Ranks may hold different frame counts (ragged tail): shards are padded to the common length for the
    equal-count collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = shard_range(n_frames, 0, world)[1]
    lo, hi = shard_range(n_frames, rank, world)
    assert local.shape[0] == hi - lo
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return out[:n_frames]


def retarget_clip_sharded(engine, raw_global_q_full, flags, ik_iters=10, damping=0.1, rot_weight=0.2, gather=True, group=None):
    """Each rank runs the fused quaternion path on its frame range of the (host-resident) clip; with
    gather=True every rank gets the whole clip's dof_pos back."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = raw_global_q_full.shape[0]
    lo, hi = shard_range(n, rank, world)
    _, dof, lp = engine.retarget_body_quat(raw_global_q_full[lo:hi], flags=flags, ik_iters=ik_iters, damping=damping,
                                           rot_weight=rot_weight, want_local_q=False)
    if gather and world > 1:
        dof = all_gather_frames(dof, n, group)
    return dof, lp
