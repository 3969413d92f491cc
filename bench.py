#!/usr/bin/env python
"""bench.py -- retargeted frames/s of the fused quaternion-path pipeline (quat mapping + angle decomposition + 10-iter IK
+ FK on a synthetic vtrdyn clip).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--total-frames T]

One "step" = one pass of the hot path over one synthetic clip.
  N = 1 : BASELINE.json configs[2], a 2^20-frame clip on one GPU.
  N > 1 : BASELINE.json configs[4] as written: a 2^24-frame clip sharded over the N GPUs (strong scaling), dof_pos
          REASSEMBLED ON EVERY RANK inside the timed region (the kernels store their dof spans straight into every
          rank's buffer over NVLink: sharding.PeerReassembly); the no-gather and the NCCL all-gather figures ride along
          as side keys.
  value  : whole-job frames/s, inputs resident in HBM, CUDA-event timed on the launching stream.
  e2e    : same metric through the reference-facing host-buffer C-ABI call (pinned host in ->
           pinned host out, H2D + D2H inside the timed region).
  roofline: algorithmic bytes (336 B in + 120 B dof + 372 B link positions = 828 B/frame,
           SURVEY.md 8(d)) / the fused kernel's CUDA-event launch time, vs the measured HBM peak.
  cpu_baseline: the CPU oracle (batched torch restatement of the reference) on a bounded sample.
--impl reference times the CPU implementation alone: the UNMODIFIED reference when its tree is present (HRT_REFERENCE or
/root/reference: dev container), else the oracle port, the travelling form of it (the GPU box has no reference tree) --
see DESIGN.md section 6.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_FRAME = 21 * 16 + 30 * 4 + 31 * 12           # 828
METRIC = "retargeted frames/s (quat mapping + angle decomposition + 10-iter IK + FK)"
IK_ITERS, DAMPING, ROT_WEIGHT = 10, 0.1, 0.2


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_capture():
    """The committed `ncu --set full` capture of this same command (profiles/traffic.json, written by tools/ncu_hot.py).
    These are PROFILER numbers of a past run, labelled as such in the line; nothing here is measured live."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p))["body_quat_kernel"]
    except Exception:
        return {}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused kernel from that capture, or None."""
    v = _ncu_capture().get("dram_bytes_per_launch")
    return float(v) if v else None


def issue_roofline(kern_ms, clocks, n_sms):
    """The bound that actually holds for the fused kernel: warp instructions issued per second against 4 schedulers per SM
    at the SM clock sampled during the run.  Instruction count per launch from the committed ncu capture."""
    n = ncu_warp_instructions()
    mhz = (clocks or {}).get("sm_mhz") or 0
    if not n or not mhz:
        return None
    peak = 4.0 * n_sms * mhz * 1e6
    achieved = n / (kern_ms * 1e-3)
    return {"bound": "issue", "achieved": achieved / 1e9, "peak": peak / 1e9, "unit": "G warp-instructions/s",
            "frac": achieved / peak, "warp_instructions_per_launch": n,
            "warp_instructions_source": "ncu capture (not live): " + str(_ncu_capture().get("source", "profiles/traffic.json"))}


def ncu_warp_instructions():
    """smsp__inst_executed.sum per launch of the fused kernel from the same committed capture, or None."""
    v = _ncu_capture().get("warp_instructions_per_launch")
    return float(v) if v else None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_samples(self, n, spin, timeout_s=6.0):
        """keep the GPU busy with `spin()` until nvidia-smi has delivered n samples"""
        t0 = time.perf_counter()
        while self.proc and len(self.rows) < n and time.perf_counter() - t0 < timeout_s:
            spin()

    def mark(self):
        return len(self.rows)

    def stop(self, first=0):
        self.rows = self.rows[first:] if first < len(self.rows) else self.rows
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_oracle_rate(frames, threads):
    """frames/s of the CPU oracle on `frames` frames of the workload (full pipeline incl. IK + FK)."""
    import torch
    from oracle import retarget_oracle as oc
    torch.set_num_threads(threads)
    sk = oc.load_skeletons()
    raw = oc.synth_clip_3q(frames, seed=0, sk=sk)
    oc.body_quat_pipeline(raw[:256], sk, clamp=True, ik_iters=IK_ITERS, damping=DAMPING, rot_weight=ROT_WEIGHT)
    t0 = time.perf_counter()
    oc.body_quat_pipeline(raw, sk, clamp=True, ik_iters=IK_ITERS, damping=DAMPING, rot_weight=ROT_WEIGHT)
    dt = time.perf_counter() - t0
    return frames / dt, dt


def reference_tree():
    """Path of the UNMODIFIED reference when it is present on this machine (dev container), else None (GPU box)."""
    p = os.environ.get("HRT_REFERENCE", "/root/reference")
    return p if os.path.isdir(os.path.join(p, "retarget", "retarget_solver")) else None


def verbatim_reference_rate(frames, threads):
    """BASELINE.md row C1q with the reference's own code: vtrdyn_zero_pose_transform batched once, then
    Mocap2HuBodyRetargeter.retarget_from_pose per frame (its per-frame Python loop; the reference has no IK stage and
    no batched form).  Returns (frames/s, seconds)."""
    import warnings
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_shim
    from make_golden import clip_3q
    torch.set_num_threads(threads)
    ref = ref_shim.load()
    warnings.simplefilter("ignore")
    src21 = ref.rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/zero_pose/vtrdyn_zero_pose.pkl"))
    tgt = ref.rkm.RobotZeroPose.from_skeleton_state(ref_shim.load_asset(ref, "asset/hu_pose/hu_v5_zero_pose.pkl"))
    raw = clip_3q(ref, frames + 50)
    solver = ref.solvers.Mocap2HuBodyRetargeter(src21, tgt)
    zq = ref.parse_mocap.vtrdyn_zero_pose_transform(raw)
    for i in range(50):
        solver.retarget_from_pose(zq[i])
    t0 = time.perf_counter()
    zq = ref.parse_mocap.vtrdyn_zero_pose_transform(raw)
    for i in range(50, frames + 50):
        solver.retarget_from_pose(zq[i])
    dt = time.perf_counter() - t0
    return frames / dt, dt


def run_reference(args):
    """The reference arm: the CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    verbatim = reference_tree() is not None and not args.force_port
    sample = min(args.ref_frames, 2000) if verbatim else args.ref_frames
    rate_fn = verbatim_reference_rate if verbatim else cpu_oracle_rate
    for _ in range(0 if verbatim else args.warmup):
        cpu_oracle_rate(min(sample, 2048), cores)
    t_all = 0.0
    steps = min(args.steps, 3) if verbatim else args.steps
    for _ in range(steps):
        r, dt = rate_fn(sample, cores)
        t_all += dt
    value = sample * steps / t_all
    if verbatim:
        kind = "reference"
        what = (f"{sample} frames/step of the configs[2] clip through the UNMODIFIED reference ({reference_tree()}): "
                "vtrdyn_zero_pose_transform batched + Mocap2HuBodyRetargeter.retarget_from_pose per frame (closed form only: the "
                f"reference has no IK / FK stage), torch threads={cores}")
    else:
        kind = "port"
        what = (f"{sample} frames/step of the 2^20-frame clip, batched torch-CPU oracle incl. {IK_ITERS}-iter IK + FK, "
                f"torch threads={torch.get_num_threads()} (no reference tree on this machine)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (+f64 Euler split)",
        "data": "synthetic",
        "config": {"workload": "configs[2]: full quaternion-path pipeline, vtrdyn (21 joints) -> Hu v5 (31 joints), "
                               f"{IK_ITERS}-iter IK + FK", "frames_per_step": sample, "ik_iters": IK_ITERS},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": what},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def side_measurements(eng, hrt, oc, sk, dev):
    """BASELINE.json configs[1] (65,536-configuration FK) and configs[3] (single-frame streaming latency of the
    teleop position path, sim_full_body_teleop.py:83-129): reported beside the headline, not part of it."""
    import numpy as np
    import torch
    out = {}
    # configs[3]: one frame at a time, pinned mapped mailboxes, host-visible in -> host-visible dof_pos
    g = torch.Generator().manual_seed(0)
    n = 100000                                             # 1000 warm-up + >= 100 k timed frames
    em = 0.4 * torch.randn(2048, 59, 3, generator=g)
    root = torch.zeros(2048, 3)
    root[:, 2] = 1.0
    _, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                      torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
    full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
    body, lh, rh = gt[:, full2body].contiguous().numpy(), gt[:, 14:34].contiguous().numpy(), gt[:, 39:59].contiguous().numpy()
    o_dof = np.empty(30, np.float32)
    lat = {"path": "VtrdynFullBodyPosRetargeter single frame, host in -> host dof_pos (hrt_stream_pos_frame via ctypes)"}
    for mode, persistent in (("launch_per_frame", False), ("resident_server", True)):
        eng.stream_pos_open(wire_layout=False, persistent=persistent)
        for i in range(1000):
            eng.stream_pos_frame(body[i % 2048], lh[i % 2048], rh[i % 2048], None, o_dof)
        ts = np.empty(n)
        for i in range(n):
            k = i % 2048
            t0 = time.perf_counter_ns()
            eng.stream_pos_frame(body[k], lh[k], rh[k], None, o_dof)
            ts[i] = time.perf_counter_ns() - t0
        paced = np.empty(240)
        period = 1.0 / 120.0
        nxt = time.perf_counter() + period
        for i in range(240):                               # 2 s at the teleop loop's 120 Hz pacing
            while time.perf_counter() < nxt:
                pass
            nxt += period
            t0 = time.perf_counter_ns()
            eng.stream_pos_frame(body[i], lh[i], rh[i], None, o_dof)
            paced[i] = time.perf_counter_ns() - t0
        eng.stream_pos_close()
        lat[mode] = {"back_to_back": {"frames": n, "p50": float(np.percentile(ts, 50)) / 1e3, "p99": float(np.percentile(ts, 99)) / 1e3,
                                      "p999": float(np.percentile(ts, 99.9)) / 1e3},
                     "paced_120hz": {"frames": 240, "p50": float(np.percentile(paced, 50)) / 1e3, "p99": float(np.percentile(paced, 99)) / 1e3}}
    # the same frames through the reference-named class, called the way sim_full_body_teleop.py:115 calls it (CPU tensors)
    solver = hrt.VtrdynFullBodyPosRetargeter(hrt.RobotZeroPose.from_asset("vtrdyn_full_zero_pose"),
                                             hrt.RobotZeroPose.from_asset("hu_v5_zero_pose"), precise_gripper=True,
                                             device=dev.index or 0)
    tb, tl, tr = torch.from_numpy(body), torch.from_numpy(lh), torch.from_numpy(rh)
    for i in range(1000):
        solver.retarget(tb[i % 2048], tl[i % 2048], tr[i % 2048], record=False)
    tc = np.empty(20000)
    for i in range(20000):
        k = i % 2048
        t0 = time.perf_counter_ns()
        solver.retarget(tb[k], tl[k], tr[k], record=False)
        tc[i] = time.perf_counter_ns() - t0
    solver._eng.stream_pos_close()
    lat["reference_class_call"] = {"call": "VtrdynFullBodyPosRetargeter.retarget(body, lhand, rhand) -> (local_q, dof_pos, body_q), CPU tensors",
                                   "frames": 20000, "p50": float(np.percentile(tc, 50)) / 1e3, "p99": float(np.percentile(tc, 99)) / 1e3}
    out["latency_us"] = lat
    # configs[1]: 65,536 Hu (33-joint) configurations, FK with joint limits; L2 flushed between launches
    eng_hu = hrt.default_engine(dev.index or 0, robot="hu")
    lo, hi = torch.tensor(oc.HU_DOF_LOWER), torch.tensor(oc.HU_DOF_UPPER)
    ang = (lo + (hi - lo) * (torch.rand(65536, 32, generator=g) * 1.2 - 0.1)).to(dev)
    gq = torch.empty((65536, 33, 4), device=dev)
    gtt = torch.empty((65536, 33, 3), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    def kernel_ms(launch, flush_l2, reps=27, skip=3):
        """Device time of ONE launch: an event pair around each of `reps` launches queued back to back (with the L2 flush in
        between where asked), one synchronize at the end: the host-side cost of issuing the call (ctypes marshalling, ~5-10 us)
        overlaps the previous kernel / the flush instead of sitting between the two events.  The event clock resolves ~1 us,
        so a 20 us launch reads 20.5, 21.5 or 22.5: the figure is the mean of the middle half of the samples (a median
        lands on one of those steps and moved by 5 % from run to run)."""
        evs = []
        torch.cuda.synchronize(dev)
        for _ in range(reps):
            if flush_l2:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            launch()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        ts = np.sort(np.array([a.elapsed_time(b) for a, b in evs[skip:]]))
        q = len(ts) // 4
        return float(ts[q:len(ts) - q].mean())

    ms = kernel_ms(lambda: eng_hu.fk_angles(hrt.TREE_ROBOT, ang, clip=True, out=(gq, gtt)), True)
    peak, _ = hbm_peak()
    jac = torch.empty((65536, 2, 6, 32), device=dev)
    jms = kernel_ms(lambda: eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, [20, 29], clip=True, out=jac), True)
    # the same FK on the host cores: the oracle's restatement of HuForwardModel.forward_kinematics (natively batched)
    ang_h = ang.cpu()
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    oc.hu_forward_kinematics(ang_h.reshape(65536, 32, 1), torch.zeros(65536, 3), oc.quat_identity((65536, 1)),
                             sk["hu_zero_pose/parents"].tolist(), torch.from_numpy(sk["hu_zero_pose/offsets"]),
                             oc.HU_DOF_AXIS, oc.HU_DOF_LOWER, oc.HU_DOF_UPPER, True)
    cpu_s = time.perf_counter() - t0
    out["fk_65536"] = {"workload": "configs[1]: Hu (33 joints) FK with limits, 65,536 configurations, L2 flushed between launches",
                       "ms": ms, "configs_per_s": 65536 / (ms * 1e-3), "algorithmic_bytes_per_config": 1080,
                       "hbm_frac": 65536 * 1080 / (ms * 1e-3) / 1e9 / peak,
                       "jacobian_2_links_ms": jms, "jacobian_hbm_frac": 65536 * 1664 / (jms * 1e-3) / 1e9 / peak,
                       "cpu_port_configs_per_s": 65536 / cpu_s,
                       "note": "device time per launch (event pair per launch, launches queued back to back); 70.8 MB = 11 us at peak, 13.6 us at the "
                               "kernel's steady-state rate: 6-8 us per launch are ramp-up and drain of a single wave (profiles/r02_notes.md)"}
    del jac, gq, gtt
    # the position-input (teleop) solver batched over a 2^18-frame clip, device-resident
    n3 = 1 << 18
    g3 = torch.Generator().manual_seed(3)
    em = 0.4 * torch.randn(n3, 59, 3, generator=g3)
    root = torch.zeros(n3, 3)
    root[:, 2] = 1.0
    _, gt3 = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                       torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
    b3, l3, r3 = gt3[:, full2body].contiguous().to(dev), gt3[:, 14:34].contiguous().to(dev), gt3[:, 39:59].contiguous().to(dev)
    d3 = torch.empty(n3, 30, device=dev)
    res = {}
    for name, fl in (("closed_form", 0), ("limits_and_10_refinement_steps", hrt.POS_CLAMP | hrt.POS_IK)):
        m3 = kernel_ms(lambda: eng.retarget_full_body_pos(b3, l3, r3, out=(None, d3, None), flags=fl), False, reps=19)
        res[name] = {"ms": m3, "frames_per_s": n3 / (m3 * 1e-3), "hbm_frac": n3 * 852 / (m3 * 1e-3) / 1e9 / peak}
    out["pos_path_2p18"] = {"workload": "config 3p: VtrdynFullBodyPosRetargeter batched, 732 B in + 120 B dof per frame, 2^18 frames (192 MB in)",
                            **res}
    return out


def synth_clip_on_device(hrt, oc, sk, dev, n, seed, chunk=1 << 20):
    """SURVEY 8(d) config-3q / config-5 recipe built on the device: local exp-maps 0.5*N(0,1) on the vtrdyn T-pose tree ->
    quaternions -> FK -> raw global quats (n,21,4).  Seeded per shard; chunked so that any rank can regenerate any chunk."""
    import torch
    parents = sk["vtrdyn_t_pose/parents"].tolist()
    off = torch.from_numpy(sk["vtrdyn_t_pose/offsets"])
    out = torch.empty((n, 21, 4), device=dev, dtype=torch.float32)
    for c, f0 in enumerate(range(0, n, chunk)):
        m = min(chunk, n - f0)
        g = torch.Generator(device=dev).manual_seed(seed * 4096 + c)
        em = 0.5 * torch.randn(m, 21, 3, device=dev, generator=g)
        lq = hrt.rotation3d.exp_map_to_quat(em)
        gq, _ = hrt.cal_forward_kinematics(lq, torch.zeros(m, 3, device=dev), parents, off, exact=True)
        out[f0:f0 + m] = gq
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import __graft_entry__ as g
    g.build()
    import humanoid_real_time_retarget_b200 as hrt
    from humanoid_real_time_retarget_b200.sharding import PeerReassembly, gpu_local_host_memory, shard_range
    from oracle import retarget_oracle as oc            # skeleton tables for input synthesis + the cpu_baseline leg only

    dev = torch.device("cuda", local_rank)
    eng = hrt.Engine(local_rank).set_standard_trees()
    flags = hrt.BQ_CLAMP | hrt.BQ_IK
    sk = oc.load_skeletons()
    multi = world > 1
    if multi:
        n_total = args.total_frames                       # configs[4]: 2^24 frames over the box, strong scaling
        lo, hi = shard_range(n_total, rank, world)
        B = hi - lo
    else:
        n_total = B = args.frames                         # configs[2]: 2^20 frames on one GPU
        lo = 0
    raw_d = synth_clip_on_device(hrt, oc, sk, dev, B, seed=1000 + rank)
    # e2e runs on a bounded host-resident sample per rank (PCIe-bound; the pinned staging buffers live on this GPU's NUMA node)
    Be = min(B, args.frames)
    with gpu_local_host_memory(dev) as numa:
        raw_h = torch.empty((Be, 21, 4), dtype=torch.float32).pin_memory()
        h_dof = torch.empty((Be, 30)).pin_memory()
        h_lp = torch.empty((Be, 31, 3)).pin_memory()
    raw_h.copy_(raw_d[:Be])
    lp_d = torch.empty((B, 31, 3), device=dev)
    dof_d = torch.empty((B, 30), device=dev)
    pr = PeerReassembly(eng, n_total, transport=args.transport) if multi else None

    def step_plain():                                      # one launch of body_quat_kernel, outputs stay on this GPU
        eng.retarget_body_quat(raw_d, flags=flags, ik_iters=IK_ITERS, damping=DAMPING, rot_weight=ROT_WEIGHT, out=(None, dof_d, lp_d))

    def step_gather():                                     # the same launch storing dof_pos into EVERY rank's clip-wide buffer + flag exchange
        pr.step(raw_d, flags, IK_ITERS, DAMPING, ROT_WEIGHT, link_pos=lp_d)

    step_dev = step_gather if multi else step_plain

    def step_e2e(dof_only=False):
        eng.retarget_body_quat_host(raw_h, flags=flags, ik_iters=IK_ITERS, damping=DAMPING, rot_weight=ROT_WEIGHT,
                                    out_dof=h_dof, out_link_pos=None if dof_only else h_lp)

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step, steps):
        """K steps between barriers: (total ms by CUDA events around the region, mean per-step ms by per-step events)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for a, b in ev:
            a.record()
            step()
            b.record()
        end.record()
        barrier()
        return start.elapsed_time(end), sum(a.elapsed_time(b) for a, b in ev) / steps

    # ---- device-resident timing (value, roofline) -------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    if rank == 0:                       # nvidia-smi needs ~0.3 s to deliver its first row: stay under load meanwhile
        sampler.wait_samples(2, lambda: (step_plain(), torch.cuda.synchronize(dev)))
    first_sample = sampler.mark()
    total_ms, kern_ms = timed(step_dev, args.steps)
    plain_total_ms = plain_kern_ms = None
    if multi:                            # side figure: the same shards with nothing communicated
        for _ in range(2):
            step_plain()
        plain_total_ms, plain_kern_ms = timed(step_plain, args.steps)
    # ---- end-to-end timing through the host-buffer C-ABI call ----------------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_dof_s = None
    if not multi:
        step_e2e(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e(True)
        barrier()
        e2e_dof_s = time.perf_counter() - t0
    # keep the same load up until a few rows have been sampled inside the measurement window
    if rank == 0:
        sampler.wait_samples(first_sample + 4, lambda: (step_plain(), torch.cuda.synchronize(dev)))
    clocks = sampler.stop(first_sample) if rank == 0 else None

    vals = [total_ms, kern_ms, e2e_s, plain_total_ms or 0.0, plain_kern_ms or 0.0]
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms, e2e_s, plain_total_ms, plain_kern_ms = t.tolist()

    # ---- N > 1: check the reassembled clip, and time the NCCL all-gather it replaces ---------------------------------
    gather_info = None
    if multi:
        step_gather()
        step_plain()
        torch.cuda.synchronize(dev)
        ok_own = bool(torch.equal(pr.dof[lo:hi], dof_d))
        # a neighbour's first chunk, regenerated here from its seed and retargeted locally, must be what that rank stored
        nb = (rank + 1) % world
        nlo, nhi = shard_range(n_total, nb, world)
        m = min(1 << 20, nhi - nlo)
        raw_nb = synth_clip_on_device(hrt, oc, sk, dev, m, seed=1000 + nb)          # = that rank's first chunk (same size, same seed)
        _, dof_nb, _ = eng.retarget_body_quat(raw_nb[:m], flags=flags, ik_iters=IK_ITERS, damping=DAMPING, rot_weight=ROT_WEIGHT,
                                              want_local_q=False, want_link_pos=False)
        ok_nb = bool(torch.equal(pr.dof[nlo:nlo + m], dof_nb))
        del raw_nb, dof_nb
        full = torch.empty((n_total, 30), device=dev)
        equal_shards = (n_total % world == 0) and (n_total // world == B)

        def step_nccl():
            step_plain()
            if equal_shards:
                dist.all_gather_into_tensor(full, dof_d)
        for _ in range(2):
            step_nccl()
        nccl_total_ms, _ = timed(step_nccl, max(3, args.steps // 2))
        nccl_ms = nccl_total_ms / max(3, args.steps // 2)
        gt = torch.tensor([nccl_ms, float(ok_own), float(ok_nb)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(gt[1:], op=dist.ReduceOp.MIN)
        nccl_ms, ok_own, ok_nb = gt.tolist()
        sent = pr.nvlink_bytes_sent_per_step
        # side figures: the other transports of the fused reassembly on the same shards
        side_ms = {}
        for other in ("packed", "multicast", "unicast"):
            if other == pr.transport or (other != "unicast" and pr.transport == "unicast"):      # unicast chosen = no NVLS here
                continue
            pr_o = PeerReassembly(eng, n_total, transport=other)
            step_o = lambda: pr_o.step(raw_d, flags, IK_ITERS, DAMPING, ROT_WEIGHT, link_pos=lp_d)  # noqa: E731
            for _ in range(2):
                step_o()
            k_o = max(3, args.steps // 2)
            o_total, _ = timed(step_o, k_o)
            ot = torch.tensor([o_total / k_o], dtype=torch.float64, device=dev)
            dist.all_reduce(ot, op=dist.ReduceOp.MAX)
            side_ms[other] = float(ot.item())
            pr_o.close()
            del pr_o
        transports = {
            "packed": "reassembly inside the compute kernel, warp-specialised: the 14 arm hinge angles of a frame (every other DOF of "
                      "this solver is structurally 0) + a 16-byte check block per 16 frames go out once through the NVSwitch multicast "
                      "address of the ranks' symmetric staging buffers (multimem.st, no flags, no fences: a group validates itself); "
                      "four unpack warps per CTA (one per scheduler, 32 registers each after setmaxnreg; the 16 compute warps keep 112) "
                      "fetch the peers' groups one round behind, check them and expand them into the local (n, 30) dof_pos under the "
                      "issue-bound solve (no NCCL on the data path)",
            "multicast": "multimem.st of full 120-byte dof rows to the NVSwitch multicast address of the ranks' symmetric buffers, issued "
                         "by the compute kernel (no NCCL on the data path)",
            "unicast": "TMA bulk stores of full 120-byte dof rows to CUDA-IPC peer buffers over NVLink, issued by the compute kernel, one "
                       "per rank and span (no NCCL on the data path)"}
        gather_info = {
            "transport": transports[pr.transport], "transport_kind": pr.transport, "multicast_unavailable": pr.transport_error,
            "packed_step_ms": side_ms.get("packed"),
            "full_row_multicast_step_ms": side_ms.get("multicast"), "full_row_unicast_step_ms": side_ms.get("unicast"),
            "payload": "dof_pos (120 B/frame) of the whole clip on every rank",
            "nvlink_bytes_sent_per_rank_per_step": sent, "nvlink_bytes_received_per_rank_per_step": pr.nvlink_bytes_received_per_step,
            "nvlink_send_GBps_per_rank": sent / (kern_ms * 1e-3) / 1e9,
            "fused_step_ms": total_ms / args.steps,
            "no_gather_step_ms": plain_total_ms / args.steps,
            "no_gather_frames_per_s": n_total * args.steps / (plain_total_ms * 1e-3),
            "nccl_compute_then_all_gather_step_ms": nccl_ms if equal_shards else None,
            "nccl_frames_per_s": (n_total / (nccl_ms * 1e-3)) if equal_shards else None,
            "reassembled_own_shard_bit_equal": bool(ok_own), "reassembled_neighbour_chunk_bit_equal": bool(ok_nb),
        }
        del full
        assert ok_own and ok_nb, "reassembled dof_pos differs from the locally computed one"

    extras = {}
    if rank == 0 and not multi and not args.no_extras:
        extras = side_measurements(eng, hrt, oc, sk, dev)
    if rank == 0:
        peak, peak_src = hbm_peak()
        value = n_total * args.steps / (total_ms * 1e-3)
        achieved = ALG_BYTES_PER_FRAME * B / (kern_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        cpu = None
        if not multi and not args.no_cpu_baseline:
            r, dt = cpu_oracle_rate(args.cpu_frames, cores)
            cpu = {"value": r, "unit": "frames/s", "cores": cores, "kind": "port",
                   "sample": f"first {args.cpu_frames} frames of the clip, batched torch-CPU oracle ({dt:.1f} s); the "
                             "reference's own per-frame Python loop runs at ~2e2 frames/s (BASELINE.md, "
                             "profiles/r01_reference_cpu_devbox.json; `bench.py --impl reference` runs it where its tree exists)"}
        if multi:
            workload = (f"configs[4]: {n_total}-frame synthetic vtrdyn clip sharded over {world} GPUs (strong scaling), full quaternion-path "
                        f"pipeline -> Hu v5, {IK_ITERS}-iter IK + FK, dof_pos reassembled on every rank INSIDE the timed region")
        else:
            workload = ("configs[2]: full quaternion-path pipeline, vtrdyn (21 joints) -> Hu v5 (31 joints), "
                        f"{IK_ITERS}-iter IK + FK")
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if multi else "weak", "vs_baseline": None, "dtype": "f32 (+f64 Euler split)", "data": "synthetic",
            "config": {"workload": workload, "frames_total_per_step": n_total, "frames_per_gpu_per_step": B, "ik_iters": IK_ITERS,
                       "l2": f"input shard {B * 336 >> 20} MB/GPU > 126 MB L2, streamed once per step", "parallelism": f"frames x{world}",
                       "e2e_frames_per_gpu_per_step": Be, "host_staging": numa.info},
            "e2e": {"value": world * Be * args.steps / e2e_s, "unit": "frames/s",
                    "h2d_bytes_per_step": Be * eng.host_input_bytes_per_frame(out_dof=True, out_link_pos=True),
                    "d2h_bytes_per_step": Be * (30 * 4 + 31 * 12)},
            "gpu_launches": args.steps * (2 if multi else 1),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "traffic_source": "ncu capture of a past run of this command, per 2^20-frame launch (not live): "
                         + str(_ncu_capture().get("source", "profiles/traffic.json")),
                         "peak_source": peak_src, "kernel": "body_quat_kernel",
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_frame": ALG_BYTES_PER_FRAME, "frames_per_launch": B,
                         "note": "with 10 IK iterations the kernel is issue-bound by construction (~0.4 Mflop/frame); "
                                 "see profiles/ for issue-slot utilisation",
                         "issue": issue_roofline(kern_ms * (1 << 20) / B, clocks, torch.cuda.get_device_properties(dev).multi_processor_count)},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if e2e_dof_s is not None:
            line["e2e_dof_only"] = {"value": Be * args.steps / e2e_dof_s, "unit": "frames/s", "h2d_bytes_per_step": Be * eng.host_input_bytes_per_frame(out_dof=True),
                                    "d2h_bytes_per_step": Be * 30 * 4,
                                    "note": "same call with out_link_pos=None: 120 B/frame back instead of 492; with the outputs the smaller side, "
                                            "only the joint range the solver reads (vtrdyn joints 10-20: 176 of 336 B per frame) crosses PCIe, "
                                            "as one strided copy per chunk"}
        line.update(extras)
        if gather_info is not None:
            line["gather"] = gather_info
        print(json.dumps(line), flush=True)
    if multi:
        pr.close()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20, help="N = 1: frames per step (configs[2]); also the e2e sample per rank")
    ap.add_argument("--total-frames", type=int, default=1 << 24, help="N > 1: frames of the whole clip per step (configs[4])")
    ap.add_argument("--transport", default="auto", choices=["auto", "packed", "multicast", "unicast"], help="N > 1: how the fused reassembly travels")
    ap.add_argument("--force-port", action="store_true", help="--impl reference: time the oracle port even where the reference tree exists")
    ap.add_argument("--cpu-frames", type=int, default=1 << 16, help="bounded CPU-baseline sample")
    ap.add_argument("--ref-frames", type=int, default=1 << 15, help="frames per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1] / configs[3] side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
