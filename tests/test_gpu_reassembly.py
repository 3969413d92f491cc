"""In-kernel reassembly (BASELINE configs[4]; hrt_retarget_body_quat_reassemble, body_quat_gather_kernel) on ONE GPU:
every rank of a 2- / 3- / 8-rank job is replayed in this process.  The peers' staging groups (14 hinge angles per frame
+ the 16-byte check block, include/hrt_b200.h) are written by the host from the single-GPU result, the rank's kernel runs
with "no multicast mapping" (d_symm_mc == d_symm: it publishes nothing) and must turn them, together with its own shard,
into the whole clip's dof_pos bit for bit.  Covers the unpack warps' fast path, ragged last groups and empty shards (the
compute warps' tail path).  The real transport (multimem.st through the NVSwitch multicast address) is checked on 2 and
8 GPUs by tools/peer_gather_check.py and bench.py (`reassembled_*_bit_equal`)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

PK_GROUP = 228          # words per 16-frame group: 224 data + 4 check


@pytest.fixture(scope="module")
def hrt():
    import __graft_entry__ as g
    g.build()
    import humanoid_real_time_retarget_b200 as h
    return h


@pytest.fixture(scope="module")
def oc():
    from oracle import retarget_oracle
    return retarget_oracle


def _salt(epoch):
    return (epoch * 0x9E3779B9 + 0x7F4A7C15) & 0xFFFFFFFF


def _pack_groups(dof, lo, n, epoch, staging):
    """Write the groups of the shard [lo, lo + n) of `dof` (numpy (N, 30)) into the staging image (uint32 view)."""
    if n == 0:
        return
    g0, groups = lo // 16, -(-n // 16)
    rows = np.zeros((groups * 16, 14), dtype=np.float32)
    shard = dof[lo:lo + n]
    rows[:n, :7] = shard[:, 11:18]                     # arm 0: Hu v5 joints 12-18 -> DOF columns 11-17
    rows[:n, 7:] = shard[:, 20:27]                     # arm 1: joints 21-27 -> columns 20-26
    words = rows.reshape(groups, 224).view(np.uint32)
    x = np.bitwise_xor.reduce(words, axis=1)
    y = words.astype(np.uint64).sum(axis=1) & 0xFFFFFFFF
    out = staging[g0 * PK_GROUP:(g0 + groups) * PK_GROUP].reshape(groups, PK_GROUP)
    out[:, :224] = words
    out[:, 224] = x ^ np.uint32(_salt(epoch))
    out[:, 225] = ((y + _salt(epoch)) & 0xFFFFFFFF).astype(np.uint32)
    out[:, 226] = epoch
    out[:, 227] = 0


@pytest.mark.parametrize("n_frames,world", [(100_003, 2), (148 * 16 * 16 * 3 + 16 * 5, 3), (1 << 18, 8), (101, 2), (7, 2)])
def test_reassembled_clip_equals_single_gpu_result(hrt, oc, n_frames, world):
    from humanoid_real_time_retarget_b200.sharding import shard_range
    sk = oc.load_skeletons()
    eng = hrt.Engine(0).set_standard_trees()
    flags = hrt.BQ_CLAMP | hrt.BQ_IK
    raw = oc.synth_clip_3q(min(n_frames, 1 << 16), seed=9, sk=sk).cuda()
    if n_frames > raw.shape[0]:
        raw = raw.repeat(-(-n_frames // raw.shape[0]), 1, 1)[:n_frames].contiguous()
    _, want, _ = eng.retarget_body_quat(raw, flags=flags, want_local_q=False, want_link_pos=False)
    want_h = want.cpu().numpy()
    shard_lo = [shard_range(n_frames, r, world)[0] for r in range(world)]
    shard_n = [shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)]
    staging_bytes, _, total_bytes, _ = eng.reassembly_layout(n_frames, shard_n)
    assert total_bytes == staging_bytes == -(-(-(-n_frames // 16) * PK_GROUP * 4) // 256) * 256
    for epoch, me in ((1, 0), (2, world - 1), (7, world // 2)):
        staging = np.zeros(total_bytes // 4, dtype=np.uint32)
        for r in range(world):
            if r != me:
                _pack_groups(want_h, shard_lo[r], shard_n[r], epoch, staging)
        symm = torch.from_numpy(staging.view(np.float32)).cuda()
        full = torch.zeros(max(n_frames, 1), 30, device="cuda")
        lo, n = shard_lo[me], shard_n[me]
        eng.retarget_body_quat_reassemble(raw[lo:lo + n] if n else raw[:0], full, n_frames, me, shard_lo, shard_n,
                                          symm.data_ptr(), symm.data_ptr(), epoch, flags=flags)
        torch.cuda.synchronize()
        assert torch.equal(full[:n_frames], want), (n_frames, world, me)


def test_stale_and_half_arrived_groups_are_refused_until_the_steps_own_land(hrt, oc):
    """The staging buffer starts with the PREVIOUS step's groups of another clip (right check sums, old epoch); this
    step's groups arrive by DMA on a second stream ~1.5 ms after the kernel has started, piecewise, while the unpack
    warps (fast path) and the compute warps' tail (ragged last group) keep fetching: nothing stale or half-arrived may
    reach the result."""
    from humanoid_real_time_retarget_b200.sharding import shard_range
    sk = oc.load_skeletons()
    eng = hrt.Engine(0).set_standard_trees()
    flags = hrt.BQ_CLAMP | hrt.BQ_IK
    n_frames, world, me, epoch = 148 * 16 * 16 * 2 + 16 * 9 + 3, 2, 0, 5
    clips = []
    for seed in (21, 22):
        raw = oc.synth_clip_3q(1 << 14, seed=seed, sk=sk).cuda().repeat(-(-n_frames // (1 << 14)), 1, 1)[:n_frames].contiguous()
        _, dof, _ = eng.retarget_body_quat(raw, flags=flags, want_local_q=False, want_link_pos=False)
        clips.append((raw, dof))
    (raw, want), (_, other) = clips
    shard_lo = [shard_range(n_frames, r, world)[0] for r in range(world)]
    shard_n = [shard_range(n_frames, r, world)[1] - shard_lo[r] for r in range(world)]
    _, _, total_bytes, _ = eng.reassembly_layout(n_frames, shard_n)
    old, new = np.zeros(total_bytes // 4, np.uint32), np.zeros(total_bytes // 4, np.uint32)
    _pack_groups(other.cpu().numpy(), shard_lo[1], shard_n[1], epoch - 1, old)
    _pack_groups(want.cpu().numpy(), shard_lo[1], shard_n[1], epoch, new)
    symm = torch.from_numpy(old.view(np.float32)).cuda()
    new_h = torch.from_numpy(new.view(np.float32)).pin_memory()
    full = torch.zeros(n_frames, 30, device="cuda")
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        torch.cuda._sleep(3_000_000)                       # ~1.5 ms: the kernel below is well under way by then
        n_piece = symm.numel() // 7 // 4 * 4
        for k in reversed(range(0, symm.numel(), n_piece)):      # the copy engine moves it in pieces, last groups first
            symm[k:k + n_piece].copy_(new_h[k:k + n_piece], non_blocking=True)
    eng.retarget_body_quat_reassemble(raw[:shard_n[0]], full, n_frames, me, shard_lo, shard_n, symm.data_ptr(), symm.data_ptr(), epoch,
                                      flags=flags)
    torch.cuda.synchronize()
    assert torch.equal(full, want)
