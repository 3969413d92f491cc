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


# ---------------------------------------------------------------------------------------------------
# Reassembly fused into the kernel's store (BASELINE configs[4]).  Every rank owns a clip-wide dof buffer; the buffers
# are exchanged once as CUDA IPC handles; from then on a step is ONE compute kernel per rank whose warps send each
# 16-frame dof span to all ranks as TMA bulk stores over NVLink, plus a one-warp flag exchange.  No NCCL call, no second
# pass over the data, and the transfer is spread over the whole kernel instead of queued behind it.
# ---------------------------------------------------------------------------------------------------
class _DevView:
    """A cudaMalloc'd span as a torch tensor (CUDA array interface); keeps nothing alive: the owner frees it."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def reassembly_transport(requested, world, multicast_available):
    """Which transport a PeerReassembly of `world` ranks ends up with.  Measured on 2^24 frames (profiles/r02_notes.md):
    full rows through the multicast address cost the solve nothing and hide under it up to 4 ranks (6.82 / 3.44 ms against
    7.07 / 3.63 ms packed); at 8 ranks their ingress, 7/8 of the clip per step, is more than one NVLink direction carries in
    the solve's time (3.21 ms against 1.94 ms packed).  Without a multicast address (no NVLS) only unicast is left."""
    if requested not in ("auto", "packed", "multicast", "unicast"):
        raise ValueError(f"unknown transport {requested!r}")
    if world <= 1 or requested == "unicast":
        return "unicast"
    if not multicast_available:
        if requested in ("packed", "multicast"):
            raise RuntimeError("multicast reassembly unavailable")
        return "unicast"
    if requested == "auto":
        return "packed" if world > 4 else "multicast"
    return requested


class PeerReassembly:
    """dof_pos of an n_frames clip, sharded by shard_range over the ranks of `group`, reassembled on every rank.
    step(engine-resident input shard) -> the clip-wide (n_frames, D) tensor of THIS rank (valid once the current stream has
    passed the step's barrier kernel).  close() must be called on every rank (it is a collective).

    transport="packed": the form the 8-GPU figure runs.  One torch symmetric-memory allocation per rank (torch.distributed
    plumbing: cuMemCreate + handle exchange + an NVSwitch multicast object over all ranks' copies) holds packed staging rows
    (the 14 arm hinge angles of a frame, 56 B instead of 120: every other DOF of this solver is structurally 0) and flags;
    the compute kernel publishes each 16-frame span once with multimem.st, raises a flag per CTA round, and unpacks the
    peers' rounds that have landed into this rank's clip-wide buffer between its own rounds (hrt_retarget_body_quat_reassemble).
    transport="multicast": full 120-byte dof rows through the multicast address, no unpacking (ingress-bound at 8 GPUs).
    transport="unicast": plain cudaMalloc buffers exchanged as CUDA IPC handles; every warp sends its span to each rank
    with one TMA bulk store per rank.
    transport="auto": with multicast (NVLS) full rows up to 4 ranks, packed beyond; without it unicast.  Full rows cost the
    solve nothing but put (N-1)/N * n_frames * 120 B on every rank's NVLink ingress per step: hidden under the solve at
    2 and 4 ranks (measured 2^24 frames: 6.81 vs 7.07 ms packed at 2, 3.88 vs 4.03 ms at 4), more than one link direction
    carries in the solve's time at 8 (3.24 vs 2.13 ms), where the packed form's unpack warps are the cheaper price."""

    def __init__(self, engine, n_frames, dof=30, group=None, transport="auto"):
        self.eng, self.group, self.n, self.D = engine, group, int(n_frames), int(dof)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 8:
            raise ValueError("peer reassembly serves the GPUs of one box (<= 8 ranks)")
        want = reassembly_transport(transport, self.world, True)       # what to try first; falls back collectively below
        self.lo, self.hi = shard_range(self.n, self.rank, self.world)
        if self.hi > self.lo and self.lo % 4:
            raise ValueError("shard boundaries must keep dof rows 16-byte aligned")
        self._buf = self._symm = None
        self.mc_ptr = 0
        self.transport_error = None
        self.packed = False
        self.shard_lo = [shard_range(self.n, r, self.world)[0] for r in range(self.world)]
        self.shard_n = [shard_range(self.n, r, self.world)[1] - shard_range(self.n, r, self.world)[0] for r in range(self.world)]
        if want != "unicast" and torch.device(engine.device).type == "cuda":
            self._try_multicast(require=(transport in ("multicast", "packed")), packed=(want == "packed"))
        self._flags, h_flags = engine.peer_alloc(64)
        if self.mc_ptr:
            self.transport = "packed" if self.packed else "multicast"
            h_buf = None
        else:
            self.transport = "unicast"
            self._buf, h_buf = engine.peer_alloc(max(self.n, 1) * self.D * 4)
        handles = [None] * self.world
        dist.all_gather_object(handles, (h_buf, h_flags), group=group)
        self.buf_ptrs, self.flag_ptrs = [], []
        for r, (hb, hf) in enumerate(handles):
            if not self.mc_ptr:
                self.buf_ptrs.append(self._buf if r == self.rank else engine.peer_open(hb))
            self.flag_ptrs.append(self._flags if r == self.rank else engine.peer_open(hf))
        self.epoch = 0
        if not self.mc_ptr:
            self.dof = torch.as_tensor(_DevView(self._buf, (self.n, self.D)), device=engine.device)
        dist.barrier(group=group)                          # every rank has mapped every buffer before the first store

    def _try_multicast(self, require, packed):
        """One symmetric allocation (the clip-wide dof buffer, or the packed staging rows + flags) with a multicast mapping;
        the decision is taken collectively (a rank that cannot map it makes every rank fall back)."""
        ok, err = 1, None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dev = torch.device(self.eng.device)
            if packed:
                _, _, total, _ = self.eng.reassembly_layout(self.n, self.shard_n)
                t = symm_mem.empty(total // 4, dtype=torch.float32, device=dev)
            else:
                t = symm_mem.empty(max(self.n, 1) * self.D, dtype=torch.float32, device=dev)
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            mc = int(hdl.multicast_ptr or 0)
            if mc == 0:
                ok, err = 0, "no multicast address (NVLS not available on this box)"
        except Exception as e:                             # no symmetric memory on this build / box
            ok, err, t, hdl, mc = 0, f"{type(e).__name__}: {e}"[:200], None, None, 0
        agree = torch.tensor([ok], dtype=torch.int32, device=self.eng.device)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=self.group)
        if int(agree.item()) == 1:
            t.zero_()
            torch.cuda.synchronize(self.eng.device)
            self._symm, self._hdl, self.mc_ptr = t, hdl, mc
            self.packed = packed
            if packed:
                self._full = torch.zeros((max(self.n, 1), self.D), dtype=torch.float32, device=self.eng.device)
                self.dof = self._full[: self.n]
                dist.barrier(group=self.group)              # every rank's staging rows and flags are zero before the first store
            else:
                self.dof = t.view(max(self.n, 1), self.D)[: self.n]
        else:
            self.transport_error = err or "another rank could not map the multicast object"
            if require:
                raise RuntimeError("multicast reassembly unavailable: " + self.transport_error)

    def step(self, raw_local, flags, ik_iters=10, damping=0.1, rot_weight=0.2, link_pos=None):
        assert raw_local.shape[0] == self.hi - self.lo
        if self.packed:
            self.epoch += 1
            self.eng.retarget_body_quat_reassemble(raw_local, self._full, self.n, self.rank, self.shard_lo, self.shard_n,
                                                   self._symm.data_ptr(), self.mc_ptr, self.epoch, flags=flags, ik_iters=ik_iters,
                                                   damping=damping, rot_weight=rot_weight, link_pos=link_pos)
            self.eng.peer_barrier(self.flag_ptrs, self.rank, self.epoch)
            return self.dof
        if self.mc_ptr:
            self.eng.retarget_body_quat_multicast(raw_local, self.mc_ptr, self.lo, flags=flags, ik_iters=ik_iters, damping=damping,
                                                  rot_weight=rot_weight, link_pos=link_pos)
        else:
            self.eng.retarget_body_quat_gather(raw_local, self.buf_ptrs, self.lo, flags=flags, ik_iters=ik_iters, damping=damping,
                                               rot_weight=rot_weight, link_pos=link_pos)
        self.epoch += 1
        self.eng.peer_barrier(self.flag_ptrs, self.rank, self.epoch)
        return self.dof

    @property
    def nvlink_bytes_sent_per_step(self):
        """Bytes this rank puts on its NVLink egress per step (multicast: its shard once; unicast: once per peer)."""
        shard = (self.hi - self.lo) * self.D * 4
        if self.world == 1:
            return 0
        if self.packed:
            return -(-(self.hi - self.lo) // 16) * 912
        return shard if self.mc_ptr else (self.world - 1) * shard

    @property
    def nvlink_bytes_received_per_step(self):
        """Bytes arriving over this rank's NVLink ingress per step (a multicast store also comes back to its sender)."""
        if self.world == 1:
            return 0
        if self.packed:
            return -(-self.n // 16) * 912
        if self.mc_ptr:
            return self.n * self.D * 4
        return (self.n - (self.hi - self.lo)) * self.D * 4

    def close(self):
        if self._flags is None:
            return
        torch.cuda.synchronize(self.eng.device)
        dist.barrier(group=self.group)                     # nobody is still storing into a buffer about to be unmapped
        for r in range(self.world):
            if r != self.rank:
                if not self.mc_ptr:
                    self.eng.peer_close(self.buf_ptrs[r])
                self.eng.peer_close(self.flag_ptrs[r])
        dist.barrier(group=self.group)
        self.dof = None
        if self._buf is not None:
            self.eng.peer_free(self._buf)
        self.eng.peer_free(self._flags)
        self._buf = self._flags = None
        self._symm = self._hdl = self._full = None         # the symmetric allocation is released with its tensor
        self.mc_ptr = 0


# ---------------------------------------------------------------------------------------------------
# Block-cyclic sharding: the all-gather of block c lands in final frame order, so it can run on a second
# stream while block c+1 is still being computed ("chunked and overlapped").
# ---------------------------------------------------------------------------------------------------
def block_cyclic_ranges(n_frames, rank, world, n_blocks):
    """Frames of the clip are cut into n_blocks super-blocks of world*blk frames; rank r owns the r-th slice of
    every super-block.  Returns (blk, [(lo, hi) per block]) with blk a multiple of 16; the clip is padded up to
    n_blocks*world*blk frames by the caller's buffers (tail ranges may be short or empty)."""
    blk = -(-n_frames // (n_blocks * world))
    blk = max(16, -(-blk // 16) * 16)
    spans = []
    for c in range(n_blocks):
        lo = min(n_frames, (c * world + rank) * blk)
        spans.append((lo, min(n_frames, lo + blk)))
    return blk, spans


def retarget_clip_overlapped(engine, raw_blocks, n_frames, flags, ik_iters=10, damping=0.1, rot_weight=0.2, n_blocks=4,
                             group=None, comm_stream=None, out=None):
    """raw_blocks: this rank's frames as a (n_blocks, blk, Js, 4) device tensor (block-cyclic slices, zero padded).
    Launches the fused kernel block by block on the current stream; as soon as a block is done its dof_pos is
    all-gathered on `comm_stream` straight into its final position of the (n_blocks*world*blk, 30) result, while
    the next block computes.  Returns the reassembled dof_pos trimmed to n_frames (valid after the returned event)."""
    world = dist.get_world_size(group)
    nb, blk = raw_blocks.shape[0], raw_blocks.shape[1]
    assert nb == n_blocks
    dev = raw_blocks.device
    if out is None:
        out = torch.empty((n_blocks * world * blk, 30), dtype=torch.float32, device=dev)
    local = torch.empty((n_blocks, blk, 30), dtype=torch.float32, device=dev)
    cuda = dev.type == "cuda"
    comm_stream = comm_stream or (torch.cuda.Stream(dev) if cuda else None)
    for c in range(n_blocks):
        engine.retarget_body_quat(raw_blocks[c], flags=flags, ik_iters=ik_iters, damping=damping, rot_weight=rot_weight,
                                  out=(None, local[c], None))
        dst = out[c * world * blk:(c + 1) * world * blk]
        if cuda:
            done = torch.cuda.Event()
            done.record()
            comm_stream.wait_event(done)
            with torch.cuda.stream(comm_stream):
                dist.all_gather_into_tensor(dst, local[c], group=group)
        else:
            dist.all_gather_into_tensor(dst, local[c], group=group)
    finished = None
    if cuda:
        finished = torch.cuda.Event()
        finished.record(comm_stream)
        torch.cuda.current_stream(dev).wait_event(finished)
    return out[:n_frames], finished


# ---------------------------------------------------------------------------------------------------
# Halo exchange: the one temporal stage after retargeting (SkeletonMotion velocities).
# np.gradient looks 1 frame either side and the sigma = 2 gaussian 8 more, so a shard needs 9 frames of each
# neighbour; with them the shard's velocities are bit-identical to the ones computed on the whole clip (same
# central differences, same fp64 filter order; the clip's two real ends keep the reference's one-sided differences
# and 'nearest' padding because rank 0 / the last rank get no halo there).
# ---------------------------------------------------------------------------------------------------
VELOCITY_HALO = 9


def exchange_halo(local, halo, group=None):
    """Neighbour exchange of `halo` leading / trailing frames between consecutive ranks (point-to-point over
    NCCL / NVLink, gloo in the CPU tests).  Returns (padded, lead): `padded` = [prev rank's tail | local | next
    rank's head], `lead` = number of frames put in front of the local ones.
    The shard lengths are agreed on first (one tiny all-gather), so that every rank takes the same decision before any
    point-to-point operation is posted: ranks without frames (ragged tails of short clips) are skipped -- a rank's
    neighbour is the nearest rank that HAS frames -- and a non-empty shard shorter than the halo raises on EVERY rank
    instead of leaving its neighbours blocked in a posted send."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    mine = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    lens_t = torch.empty(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(lens_t, mine, group=group)
    lens = lens_t.tolist()
    active = [r for r in range(world) if lens[r] > 0]
    short = [r for r in active if lens[r] < halo]
    if short and len(active) > 1:
        raise ValueError(f"shards of ranks {short} hold fewer than the {halo}-frame halo ({[lens[r] for r in short]} frames): "
                         "use fewer ranks for this clip")
    if rank not in active or len(active) == 1:
        return local, 0
    k = active.index(rank)
    prev_rank = active[k - 1] if k > 0 else None
    next_rank = active[k + 1] if k + 1 < len(active) else None
    prev_buf = torch.empty((halo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if prev_rank is not None else None
    next_buf = torch.empty((halo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) if next_rank is not None else None
    ops = []
    if prev_rank is not None:
        ops.append(dist.P2POp(dist.isend, local[:halo].contiguous(), prev_rank, group))
        ops.append(dist.P2POp(dist.irecv, prev_buf, prev_rank, group))
    if next_rank is not None:
        ops.append(dist.P2POp(dist.isend, local[-halo:].contiguous(), next_rank, group))
        ops.append(dist.P2POp(dist.irecv, next_buf, next_rank, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts = [p for p in (prev_buf, local, next_buf) if p is not None]
    return (torch.cat(parts) if len(parts) > 1 else local), (halo if prev_buf is not None else 0)


def motion_velocities_sharded(engine, local_global_t, local_global_q, dt, gaussian=True, group=None):
    """Linear and angular velocities of this rank's contiguous frame range of a clip sharded over the ranks:
    (n,J,3), (n,J,4) -> (n,J,3), (n,J,3), equal to the same frames of the whole-clip result."""
    n = local_global_t.shape[0]
    multi = dist.is_initialized() and dist.get_world_size(group) > 1
    gt, lead = exchange_halo(local_global_t, VELOCITY_HALO, group) if multi else (local_global_t, 0)
    gq, _ = exchange_halo(local_global_q, VELOCITY_HALO, group) if multi else (local_global_q, 0)
    vel = engine.motion_velocity(gt, dt, gaussian)[lead:lead + n]
    ang = engine.motion_angular_velocity(gq, dt, gaussian)[lead:lead + n]
    return vel, ang


# ---------------------------------------------------------------------------------------------------
# Host side of a rank: staging buffers next to the GPU.  The host-buffer entry points move ~830 B per frame over
# PCIe; pinned pages on the other socket cross the inter-socket link on every copy and ranks then share that link.
# ---------------------------------------------------------------------------------------------------
class gpu_local_host_memory:
    """Context manager: while active the calling thread runs on the CPUs NVML reports as closest to `device`, so
    pinned buffers allocated inside land on that GPU's NUMA node; the previous affinity is restored on exit.
    `info` says what was done ({"cpus": n, "numa_node": k} or {"skipped": reason}); never raises."""

    def __init__(self, device):
        self.device = device
        self.info = {}
        self._saved = None

    def __enter__(self):
        import os
        try:
            import pynvml
            pynvml.nvmlInit()
            props = torch.cuda.get_device_properties(self.device)
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(props.uuid)).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(torch.device(self.device).index or 0)
            self._saved = os.sched_getaffinity(0)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            now = os.sched_getaffinity(0)
            self.info = {"cpus": len(now), "of": len(self._saved)}
            try:
                self.info["numa_node"] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
            except Exception:
                pass
        except Exception as e:                                  # no NVML / no permission: keep the default placement
            self.info = {"skipped": f"{type(e).__name__}: {e}"[:120]}
        return self

    def __exit__(self, *exc):
        import os
        if self._saved is not None:
            try:
                os.sched_setaffinity(0, self._saved)
            except Exception:
                pass
        return False
