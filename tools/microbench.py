"""Per-kernel timings on the GPU box (CUDA events on the launching stream, warm-up, inputs > L2 or
L2 flushed between iterations).  Prints one JSON line per case; used to fill profiles/*.md.

    python tools/microbench.py [--cases fk,bq,elem,jac,stream] [--iters 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import __graft_entry__ as g  # noqa: E402

g.build()
import humanoid_real_time_retarget_b200 as hrt  # noqa: E402

PEAK = 6448.7
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])

_flush = None


def flush_l2():
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    _flush.zero_()


def timeit(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def report(name, n_units, bytes_per_unit, ms_med, ms_min, **extra):
    gbs = n_units * bytes_per_unit / (ms_med * 1e-3) / 1e9
    print(json.dumps({"case": name, "units": n_units, "ms_median": round(ms_med, 5), "ms_min": round(ms_min, 5),
                      "units_per_s": n_units / (ms_med * 1e-3), "alg_bytes_per_unit": bytes_per_unit,
                      "achieved_GBps": round(gbs, 1), "frac_of_hbm_peak": round(gbs / PEAK, 4), **extra}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="fk,bq,pos,elem,ops,clip,jac,stream")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--fk-log2", default="16,20,22")
    ap.add_argument("--fk-fast-only", action="store_true")
    ap.add_argument("--bq-only", default="", help="substring filter on body_quat case names")
    args = ap.parse_args()
    cases = args.cases.split(",")
    eng = hrt.default_engine(0)
    eng_hu = hrt.default_engine(0, robot="hu")
    gen = torch.Generator(device="cuda").manual_seed(0)

    if "fk" in cases:
        for L in [1 << int(x) for x in args.fk_log2.split(",")]:
            ang = (torch.rand(L, 32, device="cuda", generator=gen) - 0.5) * 2.0
            rt = torch.randn(L, 3, device="cuda", generator=gen)
            rq = torch.nn.functional.normalize(torch.randn(L, 4, device="cuda", generator=gen), dim=-1)
            out = (torch.empty(L, 33, 4, device="cuda"), torch.empty(L, 33, 3, device="cuda"))
            for exact in ((False,) if args.fk_fast_only else (False, True)):
                med, mn = timeit(lambda: eng_hu.fk_angles(hrt.TREE_ROBOT, ang, rt, rq, clip=True, exact=exact, out=out),
                                 args.iters, flush=True)
                report(f"fk_angles_hu33{'_exact' if exact else ''}", L, 156 + 924, med, mn)
            del out
        L = 1 << 20
        lq = torch.nn.functional.normalize(torch.randn(L, 33, 4, device="cuda", generator=gen), dim=-1)
        med, mn = timeit(lambda: eng_hu.fk_local_quats(hrt.TREE_ROBOT, lq), args.iters, flush=True)
        report("fk_local_quats_hu33 (incl. torch.empty)", L, 528 + 924, med, mn)

    if "bq" in cases:
        from oracle import retarget_oracle as oc            # input synthesis only
        B = 1 << 20
        raw = torch.cat([oc.synth_clip_3q(1 << 18, seed=s) for s in range(4)]).cuda()
        dof = torch.empty(B, 30, device="cuda")
        lp = torch.empty(B, 31, 3, device="cuda")
        lq = torch.empty(B, 31, 4, device="cuda")
        for name, flags, iters_ik, outs, nbytes in [
            ("bq_closed_form_dof_only", 0, 0, (None, dof, None), 336 + 120),
            ("bq_closed_form_dof+linkpos", 0, 0, (None, dof, lp), 828),
            ("bq_closed_form_all_outputs", 0, 0, (lq, dof, lp), 828 + 496),
            ("bq_clamp_dof+linkpos", hrt.BQ_CLAMP, 0, (None, dof, lp), 828),
            ("bq_ik1_dof+linkpos", hrt.BQ_CLAMP | hrt.BQ_IK, 1, (None, dof, lp), 828),
            ("bq_ik10_dof+linkpos", hrt.BQ_CLAMP | hrt.BQ_IK, 10, (None, dof, lp), 828),
        ]:
            if args.bq_only and args.bq_only not in name:
                continue
            med, mn = timeit(lambda: eng.retarget_body_quat(raw, flags=flags, ik_iters=iters_ik, out=outs), args.iters, flush=True)
            report(name, B, nbytes, med, mn)

    if "pos" in cases:
        from oracle import retarget_oracle as oc            # input synthesis only (SURVEY 8(d) config 3p recipe)
        sk = oc.load_skeletons()
        chunks = []
        for sd in range(4):
            gg = torch.Generator().manual_seed(sd)
            n = 1 << 18
            em = 0.4 * torch.randn(n, 59, 3, generator=gg)
            root = torch.zeros(n, 3)
            root[:, 2] = 1.0
            _, gt = oc.cal_forward_kinematics(oc.exp_map_to_quat(em), root, sk["vtrdyn_full_zero_pose/parents"].tolist(),
                                              torch.from_numpy(sk["vtrdyn_full_zero_pose/offsets"]))
            chunks.append(gt)
        gt = torch.cat(chunks)
        B = gt.shape[0]
        full2body = [0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 34, 35, 36, 37, 38, 39, 11, 12, 13, 14]
        body, lh, rh = gt[:, full2body].contiguous().cuda(), gt[:, 14:34].contiguous().cuda(), gt[:, 39:59].contiguous().cuda()
        dof = torch.empty(B, 30, device="cuda")
        lq = torch.empty(B, 31, 4, device="cuda")
        bq = torch.empty(B, 59, 4, device="cuda")
        med, mn = timeit(lambda: eng.retarget_full_body_pos(body, lh, rh, out=(None, dof, None)), args.iters, flush=True)
        report("pos_full_body_pos_dof_only", B, 732 + 120, med, mn)
        med, mn = timeit(lambda: eng.retarget_full_body_pos(body, lh, rh, out=(lq, dof, bq)), args.iters, flush=True)
        report("pos_full_body_pos_all_outputs (reference's 3 returns)", B, 732 + 120 + 496 + 944, med, mn)
        med, mn = timeit(lambda: eng.retarget_upper_body(body, want_local_q=False), args.iters, flush=True)
        report("pos_upper_body_dof_only (incl. torch.empty)", B, 252 + 120, med, mn)
        del body, lh, rh, dof, lq, bq, gt

    if "ops" in cases:
        from humanoid_real_time_retarget_b200 import rotation3d as r3d
        n = 1 << 24
        a = torch.nn.functional.normalize(torch.randn(n, 4, device="cuda", generator=gen), dim=-1)
        b = torch.nn.functional.normalize(torch.randn(n, 4, device="cuda", generator=gen), dim=-1)
        v = torch.randn(n, 3, device="cuda", generator=gen)
        o4, o3 = torch.empty(n, 4, device="cuda"), torch.empty(n, 3, device="cuda")
        for name, op, ins, outs, nbytes in [("op_quat_mul", r3d.OP_QUAT_MUL, [a, b], [o4], 48), ("op_quat_mul_norm", r3d.OP_QUAT_MUL_NORM, [a, b], [o4], 48),
                                            ("op_quat_rotate", r3d.OP_QUAT_ROTATE, [a, v], [o3], 40), ("op_quat_normalize", r3d.OP_QUAT_NORMALIZE, [a], [o4], 32),
                                            ("op_quat_to_exp_map", r3d.OP_QUAT_TO_EXP_MAP, [a], [o3], 28), ("op_exp_map_to_quat", r3d.OP_EXP_MAP_TO_QUAT, [v], [o4], 28)]:
            med, mn = timeit(lambda: eng.rot_op(op, n, ins, [0] * len(ins), outs), args.iters, flush=False)
            report(name, n, nbytes, med, mn)
        del a, b, v, o4, o3

    if "clip" in cases:
        T = 1 << 20
        gt = torch.randn(T, 21, 3, device="cuda", generator=gen)
        gq = torch.nn.functional.normalize(torch.randn(T, 21, 4, device="cuda", generator=gen), dim=-1)
        med, mn = timeit(lambda: eng.rescale_motion(hrt.TREE_SOURCE, gt, dir=[-1.0, -1.0, 1.0]), args.iters, flush=True)
        report("rescale_motion_21 (incl. torch.empty)", T, 504, med, mn)
        med, mn = timeit(lambda: eng.rebuild_global_rotation(hrt.TREE_SOURCE, gt), args.iters, flush=True)
        report("rebuild_global_rotation_21 (2 launches, incl. torch.empty)", T, 252 * 2 + 336, med, mn)
        med, mn = timeit(lambda: eng.motion_velocity(gt, 1 / 30), args.iters, flush=True)
        report("motion_velocity_21 (2 launches + scratch)", T, 252 * 4, med, mn)
        med, mn = timeit(lambda: eng.motion_angular_velocity(gq, 1 / 30), args.iters, flush=True)
        report("motion_angular_velocity_21 (2 launches + scratch)", T, 336 + 252 * 3, med, mn)
        del gt, gq

    if "elem" in cases:
        B = 1 << 20
        q = torch.nn.functional.normalize(torch.randn(B, 21, 4, device="cuda", generator=gen), dim=-1)
        med, mn = timeit(lambda: eng.zero_pose_transform(hrt.TREE_SOURCE, q), args.iters, flush=True)
        report("zero_pose_transform_21 (incl. torch.empty)", B, 672, med, mn)
        med, mn = timeit(lambda: eng.local_from_global(hrt.TREE_SOURCE, q), args.iters, flush=True)
        report("local_from_global_21 (incl. torch.empty)", B, 672, med, mn)

    if "jac" in cases:
        L = 1 << 18
        ang = (torch.rand(L, 32, device="cuda", generator=gen) - 0.5) * 2.0
        out = torch.empty(L, 2, 6, 32, device="cuda")
        med, mn = timeit(lambda: eng_hu.fk_jacobian(hrt.TREE_ROBOT, ang, [20, 29], clip=True, out=out), args.iters, flush=True)
        report("jacobian_hu33_K2", L, 128 + 1536, med, mn)

    if "stream" in cases:
        from oracle import retarget_oracle as oc
        raw = oc.synth_clip_3q(4096, seed=3).numpy()
        for name, flags, persistent in [("stream_closed_form", 0, False), ("stream_ik10", hrt.BQ_CLAMP | hrt.BQ_IK, False),
                                        ("stream_closed_form_resident", 0, True), ("stream_ik10_resident", hrt.BQ_CLAMP | hrt.BQ_IK, True)]:
            eng.stream_open(flags=flags, persistent=persistent)
            o_dof = np.empty(30, np.float32)
            o_lp = np.empty((31, 3), np.float32)
            for i in range(2000):
                eng.stream_frame(raw[i % 4096], None, o_dof, o_lp)
            ts = []
            for i in range(20000):
                t0 = time.perf_counter_ns()
                eng.stream_frame(raw[i % 4096], None, o_dof, o_lp)
                ts.append(time.perf_counter_ns() - t0)
            ts = np.array(ts) / 1e3
            print(json.dumps({"case": name, "frames": len(ts), "p50_us": float(np.percentile(ts, 50)),
                              "p99_us": float(np.percentile(ts, 99)), "p999_us": float(np.percentile(ts, 99.9)),
                              "max_us": float(ts.max()), "mode": "back-to-back, python caller"}), flush=True)
            eng.stream_close()


if __name__ == "__main__":
    main()
