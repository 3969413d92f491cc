// C-ABI entry points (include/hrt_b200.h) and the launch logic behind them.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/hrt_b200.h"
#include "hrt_fk.cuh"
#include "hrt_retarget.cuh"
#include "hrt_pos.cuh"
#include "hrt_ops.cuh"
#include "hrt_motion.cuh"
#include "hrt_fk_vjp.cuh"

using namespace hrt;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define HRT_CUDA(expr)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail((int)e_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

struct Tree {
    bool set = false;
    bool has_dof = false;
    TreeParams tp;
    int T = 0;                    // steps of the limb-parallel FK schedule
    bool angles_bounded = false;  // every hinge has finite limits within +-3pi/2: clipped angles need no sin/cos fallback
    float4* d_sched = nullptr;    // T * 4 entries of 32 bytes
    std::vector<float> t2z;       // host copy (J*4) or empty
    float* d_t2z = nullptr;       // device copy
    int* d_parents = nullptr;
};

constexpr int kHostStreams = 3;

}  // namespace

struct hrt_ctx {
    int device = 0;
    int sm_count = 0;
    Tree trees[HRT_MAX_TREES];
    bool bq_set = false;
    BodyQuatParams bq;
    bool pos_set[5] = {false, false, false, false, false};
    // arm hinge limits of the configured robot all finite and within +-pi: what the limit clamp / refinement needs
    // (their half-angle sine / cosine, sincos_half_lim, is written for angles inside such limits)
    bool bq_limits_ok = false;
    bool pos_limits_ok[5] = {false, false, false, false, false};
    PosParams pos[5];                 // modes 0-3 + [4] = mode 0 reading the mocap wire layout
    unsigned* d_scalars = nullptr;    // HRT_MAX_JOINTS + 2 words of device scratch for batch-wide maxima
    // staging for the *_host call
    cudaStream_t hs[kHostStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t hs_done[kHostStreams] = {nullptr, nullptr, nullptr};
    float* d_stage[kHostStreams] = {nullptr, nullptr, nullptr};
    size_t d_stage_bytes = 0;
    // streaming mailboxes (mapped pinned memory)
    bool stream_open = false;
    float* mb_in = nullptr;
    float* mb_out = nullptr;
    float *mb_in_d = nullptr, *mb_out_d = nullptr;
    cudaStream_t ss = nullptr;
    BodyQuatArgs stream_args;
    bool stream_persistent = false;    // resident server kernel for the quaternion path
    bool bserver_launched = false;
    unsigned bseq = 0;
    volatile unsigned* bctrl = nullptr;
    unsigned* bctrl_d = nullptr;
    // position-path streaming
    bool pstream_open = false;
    int pstream_mode = 0;
    bool pstream_persistent = false;   // resident server kernel instead of one launch per frame
    bool pserver_launched = false;
    unsigned pseq = 0;                 // last sequence number posted
    volatile unsigned* pctrl = nullptr; // mapped pinned control words (host view)
    unsigned* pctrl_d = nullptr;
    float *pmb_in = nullptr, *pmb_out = nullptr, *pmb_in_d = nullptr, *pmb_out_d = nullptr;
    cudaStream_t pss = nullptr;
    PosArgs pstream_args;
};

namespace {

// Every entry point runs on the context's device and puts the caller's current device back on return (a caller with
// engines on several GPUs, or torch with another current device, must not see its device change under it).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    int enter_device(int device) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return fail((int)e, "cudaGetDevice: %s", cudaGetErrorString(e));
        if (prev != device) {
            e = cudaSetDevice(device);
            if (e != cudaSuccess) return fail((int)e, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
            switched = true;
        }
        return 0;
    }
    int enter(const hrt_ctx* ctx) {
        if (!ctx) return fail(HRT_E_INVALID_ARG, "null context");
        return enter_device(ctx->device);
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};
#define HRT_ENTER(ctx)         \
    DeviceGuard dev_guard_;    \
    int rc = dev_guard_.enter(ctx); \
    if (rc) return rc

constexpr float kArmLimitMax = 3.1447f;              // 2 x the range sincos_half_lim is fitted on (a hair above pi)
constexpr unsigned long long kServerIdleNs = 20ull * 1000 * 1000;     // a resident kernel leaves after 20 ms without a frame

// Post one request to a resident server kernel and spin on its answer.  ctrl words (mapped pinned memory):
// [0] seq_in, [1] stop (host writes) / [16] seq_out, [17] exited (device writes).  `launch(served)` (re)starts the
// kernel with the last sequence number it must consider served.
template <typename Launch>
int server_roundtrip(volatile unsigned* ctrl, unsigned* seq_io, bool* launched, cudaStream_t st, Launch&& launch) {
    if (!*launched || ctrl[17] != 0u) {
        ctrl[17] = 0u;
        std::atomic_thread_fence(std::memory_order_seq_cst);
        int rc = launch(*seq_io);
        if (rc) return rc;
        *launched = true;
    }
    unsigned seq = *seq_io + 1u;
    if (seq == 0u) seq = 1u;
    std::atomic_thread_fence(std::memory_order_seq_cst);      // inputs first, then the sequence number (x86 keeps store order)
    ctrl[0] = seq;
    *seq_io = seq;
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (ctrl[16] != seq) {
        if (ctrl[17] != 0u) {                                  // the server timed out just before this request was posted
            ctrl[17] = 0u;
            std::atomic_thread_fence(std::memory_order_seq_cst);
            int rc = launch(seq - 1u);                         // it will see `seq` as new
            if (rc) return rc;
        }
        if ((++spins & 0xfffu) == 0u) {
            cudaError_t e = cudaStreamQuery(st);
            if (e != cudaSuccess && e != cudaErrorNotReady)
                return fail((int)e, "stream server failed: %s", cudaGetErrorString(e));
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2))
                return fail(HRT_E_INVALID_ARG, "stream server did not answer within 2 s");
        }
    }
    std::atomic_thread_fence(std::memory_order_seq_cst);
    return 0;
}

int get_tree(hrt_ctx* ctx, int tree, Tree** out) {
    if (tree < 0 || tree >= HRT_MAX_TREES) return fail(HRT_E_INVALID_ARG, "tree id %d out of range", tree);
    if (!ctx->trees[tree].set) return fail(HRT_E_NOT_CONFIGURED, "tree %d not installed (hrt_set_tree)", tree);
    *out = &ctx->trees[tree];
    return 0;
}

// Occupancy of (kernel, block size, dynamic shared memory) and the kernel's shared-memory attribute are queried / set once per
// device and cached per call site: both are driver calls of a few microseconds, which is a third of a 65,536-configuration
// FK launch when they sit between the caller's events.
struct LaunchCache {
    size_t occ_smem[16] = {};
    int occ_threads[16] = {};
    int per_sm[16] = {};
    size_t attr_smem[16] = {};
};

template <typename K>
int grid_for(hrt_ctx* ctx, K kernel, int threads, size_t smem, long long n_ctas_wanted, int* grid, LaunchCache* cache = nullptr) {
    const int d = ctx->device & 15;
    int per_sm = 0;
    if (cache && cache->per_sm[d] > 0 && cache->occ_smem[d] == smem && cache->occ_threads[d] == threads) {
        per_sm = cache->per_sm[d];
    } else {
        HRT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
        if (per_sm < 1) per_sm = 1;
        if (cache) { cache->per_sm[d] = per_sm; cache->occ_smem[d] = smem; cache->occ_threads[d] = threads; }
    }
    long long cap = (long long)ctx->sm_count * per_sm;
    *grid = (int)std::max(1LL, std::min(n_ctas_wanted, cap));
    return 0;
}

template <bool FROM_ANGLES, bool EXACT, bool BOUNDED>
int launch_fk_variant(hrt_ctx* ctx, Tree* t, FkArgs a, cudaStream_t st) {
    const int J = t->tp.J;
    a.sched = t->d_sched;
    a.T = t->T;
    constexpr int cfgs = FKL_GROUP * fkl_cpl(FROM_ANGLES, EXACT), warps = fkl_warps(FROM_ANGLES, EXACT);
    const size_t smem = fkl_smem_bytes(J, t->T, FROM_ANGLES, cfgs, warps);
    auto kern = fk_limb_kernel<FROM_ANGLES, EXACT, BOUNDED>;
    static LaunchCache cache;                              // one per instantiation
    const int d = ctx->device & 15;
    if (smem > 48 * 1024 && cache.attr_smem[d] < smem) {
        HRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cache.attr_smem[d] = smem;
    }
    const long long tasks = (a.B + cfgs - 1) / cfgs;
    const long long ctas = (tasks + warps - 1) / warps;
    int grid = 1;
    int rc = grid_for(ctx, kern, warps * 32, smem, ctas, &grid, &cache);
    if (rc) return rc;
    kern<<<grid, warps * 32, smem, st>>>(J, a);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

template <bool FROM_ANGLES>
int launch_fk(hrt_ctx* ctx, Tree* t, const FkArgs& a, unsigned flags, cudaStream_t st) {
    if (flags & HRT_FK_EXACT) return launch_fk_variant<FROM_ANGLES, true, false>(ctx, t, a, st);
    // clipped angles of a tree whose limits are all finite stay inside the half-angle polynomials' range
    if (FROM_ANGLES && a.clip && t->angles_bounded) return launch_fk_variant<FROM_ANGLES, false, FROM_ANGLES>(ctx, t, a, st);
    return launch_fk_variant<FROM_ANGLES, false, false>(ctx, t, a, st);
}

int body_quat_warps(const hrt_ctx* ctx, const BodyQuatArgs& a) {
    // single frames and short clips keep the small CTA; long clips: 16 warps with the refinement (issue-bound, 127 registers),
    // 28 without (the instantiation with the IK loop compiled out: latency-bound closed form, 72 registers)
    const long long groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    if (groups <= (long long)ctx->sm_count * BQ_WARPS_NARROW) return BQ_WARPS_NARROW;
    return (a.flags & BQ_IK) ? BQ_WARPS_WIDE : BQ_WARPS_NOIK;
}

size_t body_quat_smem(const hrt_ctx* ctx, const BodyQuatArgs& a);

// HRT_BQ_PACKED_IK selects the two-arms-per-thread FFMA2 variant of the refinement.  Measured on B200 it
// is ~4 % SLOWER end to end than the one-arm-per-thread kernel (profiles/r01_notes.md, "packed fp32x2"): the
// packed iteration is 17 % faster, but 168-246 registers per thread leave 8-12 warps per SM and the
// latency-bound closed-form phase loses more than the refinement gains.  Kept selectable, off by default.
bool use_packed_ik(const BodyQuatArgs& a) {
    return (a.flags & BQ_IK) && (a.flags & BQ_PACKED_IK) && !(a.flags & BQ_ACTIVE_SET) && !a.out_local_q;
}

int launch_body_quat(hrt_ctx* ctx, const BodyQuatArgs& a, cudaStream_t st, int force_grid = 0) {
    if (use_packed_ik(a)) {
        const size_t smem2 = ((size_t)BQ_CONST_WORDS + (size_t)BQ2_WARPS * bq2_warp_words(ctx->bq.J_src, ctx->bq.J_rob)) * sizeof(float);
        const long long groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
        const long long pairs = (groups + 1) / 2;
        const long long ctas = (pairs + BQ2_WARPS - 1) / BQ2_WARPS;
        const int grid = force_grid ? force_grid : (int)std::max(1LL, std::min(ctas, (long long)ctx->sm_count));
        body_quat_ik2_kernel<BQ2_WARPS><<<grid, BQ2_WARPS * 32, smem2, st>>>(ctx->bq, a);
        HRT_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = body_quat_smem(ctx, a);
    const int warps = body_quat_warps(ctx, a);
    const long long groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    const long long ctas = (groups + warps - 1) / warps;
    const int grid = force_grid ? force_grid : (int)std::max(1LL, std::min(ctas, (long long)ctx->sm_count));
    if (a.flags & BQ_IK) {
        if (warps == BQ_WARPS_NARROW) body_quat_kernel<BQ_WARPS_NARROW><<<grid, BQ_WARPS_NARROW * 32, smem, st>>>(ctx->bq, a);
        else body_quat_kernel<BQ_WARPS_WIDE><<<grid, BQ_WARPS_WIDE * 32, smem, st>>>(ctx->bq, a);
    } else {
        if (warps == BQ_WARPS_NARROW) body_quat_kernel<BQ_WARPS_NARROW, false><<<grid, BQ_WARPS_NARROW * 32, smem, st>>>(ctx->bq, a);
        else body_quat_kernel<BQ_WARPS_NOIK, false><<<grid, BQ_WARPS_NOIK * 32, smem, st>>>(ctx->bq, a);
    }
    HRT_CUDA(cudaGetLastError());
    return 0;
}

// List-schedule the tree's joints onto HRT_FK_LANES lanes: at every step each lane takes one ready
// joint (all ancestors computed in EARLIER steps), longest remaining chain first, preferring the child
// of the joint it has just computed.  Entry layout: see StepRec (32 bytes with padding).
int build_schedule(const TreeParams& tp, std::vector<float>* out, int* T_out) {
    const int J = tp.J;
    std::vector<int> height(J, 1), done(J, -1);
    for (int j = J - 1; j >= 1; --j) height[tp.parent[j]] = std::max(height[tp.parent[j]], height[j] + 1);
    done[0] = -1;                                   // the root is available before step 0
    std::vector<char> scheduled(J, 0);
    scheduled[0] = 1;
    int remaining = J - 1, T = 0;
    int last[HRT_FK_LANES];
    for (int p = 0; p < HRT_FK_LANES; ++p) last[p] = -1;
    out->clear();
    while (remaining > 0) {
        if (T >= HRT_MAX_STEPS) return -1;
        int pick[HRT_FK_LANES];
        for (int p = 0; p < HRT_FK_LANES; ++p) pick[p] = -1;
        std::vector<char> taken(J, 0);
        // first pass: continue a chain (child of the joint this lane computed in the previous step)
        for (int pass = 0; pass < 2; ++pass)
            for (int p = 0; p < HRT_FK_LANES; ++p) {
                if (pick[p] >= 0) continue;
                int best = -1;
                for (int j = 1; j < J; ++j) {
                    if (scheduled[j] || taken[j]) continue;
                    const int par = tp.parent[j];
                    if (!scheduled[par] || done[par] >= T) continue;          // parent not finished before this step
                    if (pass == 0 && par != last[p]) continue;
                    if (best < 0 || height[j] > height[best]) best = j;
                }
                if (best >= 0) { pick[p] = best; taken[best] = 1; }
            }
        for (int p = 0; p < HRT_FK_LANES; ++p) {
            const int j = pick[p];
            // 32-byte record: offset xyz | w3 || limits | w6 | w7, the w words carrying the BYTE offsets the kernel adds to
            // its row bases: w3 = live << 31 | axis << 16 | (j - 1) * 4  (angle row), w6 = par * 16 | par * 12 << 16 (parent in the
            // quaternion / position image), w7 = j * 16 | j * 12 << 16 (this joint's slots).  An idle lane-step reads joint 1
            // and parent 0 and stores nothing.
            float rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int jj = j >= 0 ? j : 1, par = j >= 0 ? tp.parent[j] : 0;
            uint32_t w3 = (uint32_t)((jj - 1) * 4) | ((uint32_t)jr_axis(tp.jr[jj].meta) << 16);
            const uint32_t w6 = (uint32_t)(par * 16) | ((uint32_t)(par * 12) << 16);
            const uint32_t w7 = (uint32_t)(jj * 16) | ((uint32_t)(jj * 12) << 16);
            if (j >= 0) {
                w3 |= 0x80000000u;
                rec[0] = tp.jr[j].off[0]; rec[1] = tp.jr[j].off[1]; rec[2] = tp.jr[j].off[2];
                rec[4] = tp.lim[j][0]; rec[5] = tp.lim[j][1];
                scheduled[j] = 1;
                done[j] = T;
                --remaining;
            }
            memcpy(&rec[3], &w3, 4);
            memcpy(&rec[6], &w6, 4);
            memcpy(&rec[7], &w7, 4);
            out->insert(out->end(), rec, rec + 8);
            last[p] = j;
        }
        ++T;
    }
    *T_out = T;
    return 0;
}

int fill_body_quat_args(hrt_ctx* ctx, int64_t B, const float* src, unsigned flags, int ik_iters, float damping,
                        float rot_weight, float* lq, float* dof, float* lp, BodyQuatArgs* a) {
    if (!ctx->bq_set) return fail(HRT_E_NOT_CONFIGURED, "hrt_configure_body_quat has not been called");
    if (B < 0) return fail(HRT_E_INVALID_ARG, "negative frame count");
    if ((flags & HRT_BQ_IK) && (ik_iters < 0 || ik_iters > 1000)) return fail(HRT_E_INVALID_ARG, "ik_iters out of range");
    if ((flags & (HRT_BQ_CLAMP | HRT_BQ_IK)) && !ctx->bq_limits_ok)
        return fail(HRT_E_UNSUPPORTED_TREE, "limits / refinement need finite arm hinge limits that do not reach beyond +-pi in the robot tree");
    a->B = B;
    a->src_gq = src;
    a->pre_transformed = (flags & HRT_BQ_PRE_TRANSFORMED) ? 1 : 0;
    a->flags = (flags & HRT_BQ_CLAMP ? BQ_CLAMP : 0u) | (flags & HRT_BQ_IK ? BQ_IK : 0u) | (flags & HRT_BQ_PACKED_IK ? BQ_PACKED_IK : 0u) |
               (flags & HRT_BQ_ACTIVE_SET ? BQ_ACTIVE_SET : 0u);
    a->ik_iters = ik_iters;
    a->damping = damping;
    a->rot_weight = rot_weight;
    a->out_local_q = lq;
    a->out_dof = dof;
    a->out_link_pos = lp;
    a->n_peer = 0;
    a->peer_frame0 = 0;
    for (int r = 0; r < HRT_MAX_PEERS; ++r) a->peer_dof[r] = nullptr;
    a->mc_dof = nullptr;
    memset(&a->g, 0, sizeof(a->g));
    return 0;
}

}  // namespace

namespace {
size_t body_quat_smem(const hrt_ctx* ctx, const BodyQuatArgs& a) {
    return ((size_t)BQ_CONST_WORDS +
            (size_t)body_quat_warps(ctx, a) * bq_tile_words(ctx->bq.J_src, ctx->bq.J_rob, a.out_local_q != nullptr)) * sizeof(float);
}
}  // namespace

extern "C" {

int hrt_abi_version(void) { return HRT_ABI_VERSION; }

const char* hrt_last_error_string(void) { return g_err; }

int hrt_ctx_create(int device, hrt_ctx** out) {
    if (!out) return fail(HRT_E_INVALID_ARG, "out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(HRT_E_NO_DEVICE, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(HRT_E_INVALID_ARG, "device %d out of range (0..%d)", device, n - 1);
    DeviceGuard dev_guard_;
    {
        int rc = dev_guard_.enter_device(device);
        if (rc) return rc;
    }
    hrt_ctx* c = new (std::nothrow) hrt_ctx();
    if (!c) return fail(HRT_E_INVALID_ARG, "out of host memory");
    c->device = device;
    HRT_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    // these kernels need more than the default 48 KB of dynamic shared memory
    HRT_CUDA(cudaFuncSetAttribute(jacobian_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(jacobian_kernel<HRT_MAX_CHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(body_quat_kernel<BQ_WARPS_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(body_quat_kernel<BQ_WARPS_NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(body_quat_kernel<BQ_WARPS_NOIK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(body_quat_kernel<BQ_WARPS_NARROW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(body_quat_ik2_kernel<BQ2_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
#define HRT_POS_ATTR(MODE)                                                                                                          \
    HRT_CUDA(cudaFuncSetAttribute(pos_retarget_kernel<MODE, POS_WARPS_MIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    HRT_CUDA(cudaFuncSetAttribute(pos_retarget_kernel<MODE, POS_WARPS_MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); \
    HRT_CUDA(cudaFuncSetAttribute(pos_retarget_kernel<MODE, POS_WARPS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    HRT_POS_ATTR(POS_FULL_BODY_POS)
    HRT_POS_ATTR(POS_UPPER_BODY)
    HRT_POS_ATTR(POS_FULL_BODY)
    HRT_POS_ATTR(POS_MAIN)
#undef HRT_POS_ATTR
    HRT_CUDA(cudaFuncSetAttribute(rescale_motion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HRT_CUDA(cudaFuncSetAttribute(rebuild_rotation_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    HRT_CUDA(cudaMalloc(&c->d_scalars, (HRT_MAX_JOINTS + 2) * sizeof(unsigned)));
    *out = c;
    return 0;
}

int hrt_ctx_destroy(hrt_ctx* ctx) {
    if (!ctx) return 0;
    DeviceGuard dev_guard_;
    dev_guard_.enter(ctx);
    hrt_stream_close(ctx);
    hrt_stream_pos_close(ctx);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    for (auto& t : ctx->trees) {
        if (t.d_t2z) cudaFree(t.d_t2z);
        if (t.d_parents) cudaFree(t.d_parents);
        if (t.d_sched) cudaFree(t.d_sched);
    }
    for (int i = 0; i < kHostStreams; ++i) {
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
        if (ctx->hs_done[i]) cudaEventDestroy(ctx->hs_done[i]);
        if (ctx->hs[i]) cudaStreamDestroy(ctx->hs[i]);
    }
    delete ctx;
    return 0;
}

int hrt_ctx_sm_count(const hrt_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int hrt_set_tree(hrt_ctx* ctx, int tree, int J, const int32_t* parents, const float* offsets,
                 const uint8_t* dof_axis, const float* lower, const float* upper, const float* t2z) {
    HRT_ENTER(ctx);
    // streams (and a resident server kernel) hold the previous tables by value: stop them, the caller re-opens
    hrt_stream_close(ctx);
    hrt_stream_pos_close(ctx);
    if (tree < 0 || tree >= HRT_MAX_TREES) return fail(HRT_E_INVALID_ARG, "tree id %d out of range", tree);
    if (J < 1 || J > HRT_MAX_JOINTS) return fail(HRT_E_UNSUPPORTED_TREE, "J=%d outside 1..%d", J, HRT_MAX_JOINTS);
    if (!parents || !offsets) return fail(HRT_E_INVALID_ARG, "parents/offsets are required");
    if (parents[0] != -1) return fail(HRT_E_UNSUPPORTED_TREE, "joint 0 must be the root (parent -1)");
    for (int j = 1; j < J; ++j)
        if (parents[j] < 0 || parents[j] >= j)
            return fail(HRT_E_UNSUPPORTED_TREE, "parents[%d]=%d: parents must precede children, single root", j, parents[j]);
    Tree& t = ctx->trees[tree];
    t.set = false;                     // an early return below must not leave a half-installed tree usable
    TreeParams& tp = t.tp;
    memset(&tp, 0, sizeof(tp));
    tp.J = J;
    // liveness-based slot allocation for parents that are not the previous joint
    std::vector<int> last_use(J, -1);
    for (int j = 1; j < J; ++j)
        if (parents[j] != j - 1) last_use[parents[j]] = j;
    int owner[HRT_MAX_SLOTS];
    for (int s = 0; s < HRT_MAX_SLOTS; ++s) owner[s] = -1;
    std::vector<int> slot_of(J, -1);
    int n_slots = 0;
    if (dof_axis)
        for (int j = 1; j < J; ++j)
            if (dof_axis[j - 1] > 2) return fail(HRT_E_INVALID_ARG, "dof_axis[%d]=%d not in 0..2", j - 1, dof_axis[j - 1]);
    for (int j = 0; j < J; ++j) {
        tp.parent[j] = (int8_t)parents[j];
        int src = -1, save = -1;
        if (j > 0 && parents[j] != j - 1) src = slot_of[parents[j]];
        if (last_use[j] > j) {
            int s = -1;
            for (int k = 0; k < HRT_MAX_SLOTS; ++k)
                if (owner[k] < 0 || last_use[owner[k]] <= j) { s = k; break; }
            if (s < 0) return fail(HRT_E_UNSUPPORTED_TREE, "tree needs more than %d live branch points", HRT_MAX_SLOTS);
            owner[s] = j;
            slot_of[j] = s;
            save = s;
            n_slots = std::max(n_slots, s + 1);
        }
        for (int k = 0; k < 3; ++k) tp.jr[j].off[k] = offsets[j * 3 + k];
        const uint32_t axis = (dof_axis && j > 0) ? dof_axis[j - 1] : 0u;
        tp.jr[j].meta = axis | ((uint32_t)(src + 1) << 4) | ((uint32_t)(save + 1) << 8);
        tp.lim[j][0] = (lower && j > 0) ? lower[j - 1] : -INFINITY;
        tp.lim[j][1] = (upper && j > 0) ? upper[j - 1] : INFINITY;
    }
    tp.n_slots = n_slots;
    t.has_dof = dof_axis != nullptr;
    t.angles_bounded = lower && upper;
    for (int j = 1; j < J && t.angles_bounded; ++j)
        t.angles_bounded = std::fabs(tp.lim[j][0]) <= 4.7f && std::fabs(tp.lim[j][1]) <= 4.7f;      // NaN fails too
    if (t.d_t2z) { cudaFree(t.d_t2z); t.d_t2z = nullptr; }
    if (t.d_parents) { cudaFree(t.d_parents); t.d_parents = nullptr; }
    if (t.d_sched) { cudaFree(t.d_sched); t.d_sched = nullptr; }
    {
        std::vector<float> sched;
        if (build_schedule(tp, &sched, &t.T)) return fail(HRT_E_UNSUPPORTED_TREE, "FK schedule longer than %d steps", HRT_MAX_STEPS);
        if (!sched.empty()) {
            HRT_CUDA(cudaMalloc(&t.d_sched, sched.size() * sizeof(float)));
            HRT_CUDA(cudaMemcpy(t.d_sched, sched.data(), sched.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    t.t2z.clear();
    HRT_CUDA(cudaMalloc(&t.d_parents, J * sizeof(int)));
    HRT_CUDA(cudaMemcpy(t.d_parents, parents, J * sizeof(int), cudaMemcpyHostToDevice));
    if (t2z) {
        t.t2z.assign(t2z, t2z + J * 4);
        HRT_CUDA(cudaMalloc(&t.d_t2z, J * 4 * sizeof(float)));
        HRT_CUDA(cudaMemcpy(t.d_t2z, t2z, J * 4 * sizeof(float), cudaMemcpyHostToDevice));
    }
    t.set = true;
    return 0;
}

int hrt_fk_local_quats(hrt_ctx* ctx, int tree, int64_t B, const float* d_local_q, const float* d_root_t,
                       float* d_gq, float* d_gt, unsigned flags, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (B == 0) return 0;
    if (B < 0 || !d_local_q) return fail(HRT_E_INVALID_ARG, "bad B / null input");
    if (!aligned16(d_local_q) || !aligned16(d_gq) || !aligned16(d_gt) || !aligned16(d_root_t))
        return fail(HRT_E_ALIGNMENT, "device buffers must be 16-byte aligned");
    FkArgs a{};
    a.B = B; a.local_q = d_local_q; a.root_t = d_root_t; a.out_gq = d_gq; a.out_gt = d_gt;
    return launch_fk<false>(ctx, t, a, flags, (cudaStream_t)stream);
}

int hrt_fk_angles(hrt_ctx* ctx, int tree, int64_t B, const float* d_angles, const float* d_root_t,
                  const float* d_root_q, int clip, float* d_gq, float* d_gt, unsigned flags, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (!t->has_dof) return fail(HRT_E_NOT_CONFIGURED, "tree %d has no dof_axis table", tree);
    if (B == 0) return 0;
    if (B < 0 || !d_angles) return fail(HRT_E_INVALID_ARG, "bad B / null input");
    if (!aligned16(d_root_q) || !aligned16(d_gq) || !aligned16(d_angles) || !aligned16(d_gt) || !aligned16(d_root_t))
        return fail(HRT_E_ALIGNMENT, "device buffers must be 16-byte aligned");
    FkArgs a{};
    a.B = B; a.angles = d_angles; a.root_t = d_root_t; a.root_q = d_root_q; a.out_gq = d_gq; a.out_gt = d_gt;
    a.clip = clip;
    return launch_fk<true>(ctx, t, a, flags, (cudaStream_t)stream);
}

int hrt_fk_jacobian(hrt_ctx* ctx, int tree, int64_t B, const float* d_angles, const float* d_root_t,
                    const float* d_root_q, int clip, const int32_t* links, int K, float* d_jac, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (!t->has_dof) return fail(HRT_E_NOT_CONFIGURED, "tree %d has no dof_axis table", tree);
    if (B < 0 || !links || (B > 0 && (!d_angles || !d_jac))) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    if (K < 1 || K > HRT_MAX_LINKS) return fail(HRT_E_INVALID_ARG, "K=%d outside 1..%d", K, HRT_MAX_LINKS);
    if (!aligned16(d_root_q)) return fail(HRT_E_ALIGNMENT, "root_q must be 16-byte aligned");
    JacParams jp;
    memset(&jp, 0, sizeof(jp));
    jp.K = K;
    for (int k = 0; k < K; ++k) {
        int l = links[k];
        if (l < 0 || l >= t->tp.J) return fail(HRT_E_INVALID_ARG, "links[%d]=%d out of range", k, l);
        int tmp[HRT_MAX_JOINTS], n = 0;
        for (int j = l; j > 0; j = t->tp.parent[j]) tmp[n++] = j;
        if (n > HRT_MAX_CHAIN) return fail(HRT_E_UNSUPPORTED_TREE, "chain to link %d deeper than %d", l, HRT_MAX_CHAIN);
        jp.link[k] = l;
        jp.depth[k] = n;
        for (int c = 0; c < n; ++c) jp.chain[k][c] = (int8_t)tmp[n - 1 - c];
    }
    if (B == 0) return 0;
    FkArgs a{};
    a.B = B; a.angles = d_angles; a.root_t = d_root_t; a.root_q = d_root_q; a.out_jac = d_jac; a.clip = clip;
    const int D = t->tp.J - 1;
    if ((6 * D) % 4 != 0) return fail(HRT_E_UNSUPPORTED_TREE, "the Jacobian kernel needs 6*(J-1) to be a multiple of 4");
    if (!aligned16(d_jac)) return fail(HRT_E_ALIGNMENT, "d_jac must be 16-byte aligned");
    const int cpw = 32 / jac_lanes_per_cfg(K);
    const size_t smem = (size_t)JAC_WARPS_PER_CTA * jac_tile_words(K, D, cpw) * sizeof(float);
    const long long groups = (B + cpw - 1) / cpw;
    const long long ctas = (groups + JAC_WARPS_PER_CTA - 1) / JAC_WARPS_PER_CTA;
    int grid = 1;
    int max_depth = 0;
    for (int k = 0; k < K; ++k) max_depth = std::max(max_depth, jp.depth[k]);
    static LaunchCache jc8, jc16;
    if (max_depth <= 8) {
        if ((rc = grid_for(ctx, jacobian_kernel<8>, JAC_WARPS_PER_CTA * 32, smem, ctas, &grid, &jc8))) return rc;
        jacobian_kernel<8><<<grid, JAC_WARPS_PER_CTA * 32, smem, (cudaStream_t)stream>>>(t->tp, jp, a);
    } else {
        if ((rc = grid_for(ctx, jacobian_kernel<HRT_MAX_CHAIN>, JAC_WARPS_PER_CTA * 32, smem, ctas, &grid, &jc16))) return rc;
        jacobian_kernel<HRT_MAX_CHAIN><<<grid, JAC_WARPS_PER_CTA * 32, smem, (cudaStream_t)stream>>>(t->tp, jp, a);
    }
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_fk_vjp(hrt_ctx* ctx, int tree, int64_t B, const float* d_angles, const float* d_root_t, const float* d_root_q, int clip,
               const float* d_g_gq, const float* d_g_gt, float* d_g_angles, float* d_g_root_t, float* d_g_root_q, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (!t->has_dof) return fail(HRT_E_NOT_CONFIGURED, "tree %d has no dof_axis table", tree);
    if (B == 0) return 0;
    if (B < 0 || !d_angles || !d_g_angles) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    if (!aligned16(d_root_q) || !aligned16(d_g_gq) || !aligned16(d_g_root_q))
        return fail(HRT_E_ALIGNMENT, "root_q, g_gq and g_root_q must be 16-byte aligned");
    FkVjpArgs a{};
    a.B = B; a.angles = d_angles; a.root_t = d_root_t; a.root_q = d_root_q; a.clip = clip;
    a.g_gq = d_g_gq; a.g_gt = d_g_gt; a.g_angles = d_g_angles; a.g_root_t = d_g_root_t; a.g_root_q = d_g_root_q;
    const size_t smem = vjp_smem_bytes(t->tp.J - 1);
    const long long groups = (B + 31) / 32;
    int grid = 1;
    static LaunchCache cache;
    if ((rc = grid_for(ctx, fk_vjp_kernel, VJP_WARPS * 32, smem, (groups + VJP_WARPS - 1) / VJP_WARPS, &grid, &cache))) return rc;
    fk_vjp_kernel<<<grid, VJP_WARPS * 32, smem, (cudaStream_t)stream>>>(t->tp, a);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_ik_refine(hrt_ctx* ctx, int64_t B, const float* d_theta0, const float* d_pe_t, const float* d_pw_t, const float* d_qw_t,
                  int iters, float damping, float rot_weight, unsigned flags, float* d_theta, float* d_residual, void* stream) {
    HRT_ENTER(ctx);
    if (!ctx->bq_set) return fail(HRT_E_NOT_CONFIGURED, "hrt_configure_body_quat has not been called (it installs the robot's arm tables)");
    if (!ctx->bq_limits_ok)
        return fail(HRT_E_UNSUPPORTED_TREE, "limits / refinement need finite arm hinge limits that do not reach beyond +-pi in the robot tree");
    if (iters < 0 || iters > 1000) return fail(HRT_E_INVALID_ARG, "iters out of range");
    if (B == 0) return 0;
    if (B < 0 || !d_theta0 || !d_pe_t || !d_pw_t || !d_qw_t || !d_theta) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    if (!aligned16(d_qw_t)) return fail(HRT_E_ALIGNMENT, "qw_t must be 16-byte aligned");
    IkArmTables tb;
    for (int side = 0; side < 2; ++side) {
        const ArmParams& ap = ctx->bq.arm[side];
        memcpy(tb.off[side], ap.off, sizeof(ap.off));
        memcpy(tb.lower[side], ap.lower, sizeof(ap.lower));
        memcpy(tb.upper[side], ap.upper, sizeof(ap.upper));
        for (int k = 0; k < 3; ++k) tb.p_sh[side][k] = ctx->bq.shoulder_p[side][k];
    }
    IkRefineArgs a{};
    a.B = B; a.theta0 = d_theta0; a.pe_t = d_pe_t; a.pw_t = d_pw_t; a.qw_t = d_qw_t; a.theta = d_theta; a.residual = d_residual;
    a.iters = iters; a.damping = damping; a.rot_weight = rot_weight; a.active_set = (flags & HRT_BQ_ACTIVE_SET) ? 1 : 0;
    const int grid = (int)std::max(1LL, std::min((2 * (long long)B + 127) / 128, (long long)ctx->sm_count * 8));
    ik_refine_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(tb, a);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_local_from_global(hrt_ctx* ctx, int tree, int64_t B, const float* d_gq, float* d_lq, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (B == 0) return 0;
    if (B < 0 || !d_gq || !d_lq) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    if (!aligned16(d_gq) || !aligned16(d_lq)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    const long long n = (long long)B * t->tp.J;
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)ctx->sm_count * 16);
    local_from_global_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_gq), t->d_parents, t->tp.J, n, reinterpret_cast<float4*>(d_lq));
    HRT_CUDA(cudaGetLastError());
    return 0;
}

static void rot_quat(int variant, float out[4]) {
    // quat_from_angle_axis(torch.tensor(pi/2), axis) as the reference evaluates it in fp32
    // (rotation3d.py:123-143): sin(pi/4) = cos(pi/4) = 0x3f3504f3, norm 0x3f7fffff, quotient
    // 0x3f3504f4 (= 0.70710683).  Pinned by tests/golden/zero_pose_transform.npz.
    union { uint32_t u; float f; } h;
    h.u = 0x3f3504f4u;
    out[0] = out[1] = out[2] = 0.f;
    out[variant == 1 ? 0 : 2] = h.f;
    out[3] = h.f;
}

int hrt_zero_pose_transform(hrt_ctx* ctx, int tree, int64_t B, const float* d_gq, int variant, float* d_out,
                            void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (!t->d_t2z) return fail(HRT_E_NOT_CONFIGURED, "tree %d has no T2Z table", tree);
    if (variant != 0 && variant != 1) return fail(HRT_E_INVALID_ARG, "variant must be 0 (z) or 1 (x, broadcast)");
    if (B == 0) return 0;
    if (B < 0 || !d_gq || !d_out) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    if (!aligned16(d_gq) || !aligned16(d_out)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    float r[4];
    rot_quat(variant, r);
    const long long n = (long long)B * t->tp.J;
    // the CTAs that are resident together (a second, partly filled wave of a grid-stride kernel is a tail)
    static LaunchCache zc;
    int grid = 1;
    if ((rc = grid_for(ctx, zero_pose_transform_kernel, 256, 0, (n + 256 * ZPT_UNROLL - 1) / (256 * ZPT_UNROLL), &grid, &zc))) return rc;
    zero_pose_transform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(d_gq), reinterpret_cast<const float4*>(t->d_t2z),
        make_float4(r[0], r[1], r[2], r[3]), t->tp.J, n, reinterpret_cast<float4*>(d_out));
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_configure_body_quat(hrt_ctx* ctx, int src_tree, int rob_tree, const int32_t* src_joints,
                            const int32_t* rob_first) {
    HRT_ENTER(ctx);
    hrt_stream_close(ctx);             // a stream opened before keeps the old wiring by value
    Tree *s, *r;
    if ((rc = get_tree(ctx, src_tree, &s))) return rc;
    if ((rc = get_tree(ctx, rob_tree, &r))) return rc;
    if (!src_joints || !rob_first) return fail(HRT_E_INVALID_ARG, "null joint tables");
    if (s->t2z.empty()) return fail(HRT_E_NOT_CONFIGURED, "source tree has no T2Z table");
    if (!r->has_dof) return fail(HRT_E_NOT_CONFIGURED, "robot tree has no dof tables");
    const TreeParams& rt = r->tp;
    BodyQuatParams& bp = ctx->bq;
    memset(&bp, 0, sizeof(bp));
    bp.J_src = s->tp.J;
    bp.J_rob = rt.J;
    if (bp.J_rob > 31 || bp.J_src > 31)
        return fail(HRT_E_UNSUPPORTED_TREE, "skeleton too large for the fused kernel tile (max 31 joints)");
    rot_quat(0, bp.rot_z90);
    // zero-pose positions of every robot joint: p_j = off_j + p_parent (identity rotations)
    std::vector<float> pos(rt.J * 3, 0.f);
    for (int j = 1; j < rt.J; ++j)
        for (int k = 0; k < 3; ++k) pos[j * 3 + k] = rt.jr[j].off[k] + pos[rt.parent[j] * 3 + k];
    for (int i = 0; i < rt.J * 3; ++i) bp.rest_pos[i] = pos[i];
    for (int j = 1; j < rt.J; ++j) {
        bp.lower_all[j] = rt.lim[j][0];
        bp.upper_all[j] = rt.lim[j][1];
        bp.axis_all[j] = (uint8_t)jr_axis(rt.jr[j].meta);
    }
    static const int arm_axis[7] = {1, 0, 2, 1, 0, 1, 2};
    bool limits_ok = true;
    for (int side = 0; side < 2; ++side) {
        ArmParams& ap = bp.arm[side];
        const int32_t* sj = src_joints + side * 5;
        for (int n = 0; n < 5; ++n)
            if (sj[n] < 0 || sj[n] >= s->tp.J) return fail(HRT_E_INVALID_ARG, "src_joints[%d][%d]=%d out of range", side, n, sj[n]);
        // the fused kernel takes the upper / lower arm's local rotation against the listed shoulder / upper arm joint:
        // that is cal_local_rotation (kinematics.py:41-63) only if those ARE the tree parents
        if (s->tp.parent[sj[2]] != sj[1] || s->tp.parent[sj[3]] != sj[2])
            return fail(HRT_E_UNSUPPORTED_TREE, "src_joints[%d]: joint %d must be the parent of %d and %d the parent of %d in the source tree",
                        side, sj[1], sj[2], sj[2], sj[3]);
        ap.src_torso = sj[0]; ap.src_shoulder = sj[1]; ap.src_upper = sj[2]; ap.src_lower = sj[3]; ap.src_hand = sj[4];
        const int f = rob_first[side];
        if (f < 1 || f + 8 >= rt.J) return fail(HRT_E_INVALID_ARG, "rob_first[%d]=%d out of range", side, f);
        ap.rob_first = f;
        for (int c = 0; c < 7; ++c) {
            if (jr_axis(rt.jr[f + c].meta) != arm_axis[c])
                return fail(HRT_E_UNSUPPORTED_TREE, "arm hinge %d has axis %d, the fused kernel is built for (y,x,z,y,x,y,z)", c,
                            jr_axis(rt.jr[f + c].meta));
            if (c > 0 && rt.parent[f + c] != f + c - 1) return fail(HRT_E_UNSUPPORTED_TREE, "arm joints must form a chain");
            ap.lower[c] = rt.lim[f + c][0];
            ap.upper[c] = rt.lim[f + c][1];
            if (!(ap.lower[c] >= -kArmLimitMax && ap.upper[c] <= kArmLimitMax)) limits_ok = false;
        }
        if (rt.parent[f + 7] != f + 6 || rt.parent[f + 8] != f + 6)
            return fail(HRT_E_UNSUPPORTED_TREE, "expected two gripper links under the wrist-yaw link");
        for (int c = 0; c < 9; ++c)
            for (int k = 0; k < 3; ++k) ap.off[c][k] = rt.jr[f + c].off[k];
        for (int n = 0; n < 5; ++n)
            for (int k = 0; k < 4; ++k) ap.t2z[n][k] = s->t2z[sj[n] * 4 + k];
        for (int k = 0; k < 3; ++k) {
            bp.shoulder_p[side][k] = pos[f * 3 + k];
            ap.seg_elbow[k] = pos[(f + 3) * 3 + k] - pos[f * 3 + k];
            ap.seg_wrist[k] = pos[(f + 6) * 3 + k] - pos[(f + 3) * 3 + k];
        }
        // the arm hangs off the torso link, which keeps identity rotation in this pipeline
        const int torso = rt.parent[f];
        for (int k = 0; k < 3; ++k) bp.torso_p[k] = pos[torso * 3 + k];
    }
    ctx->bq_set = true;
    ctx->bq_limits_ok = limits_ok;
    return 0;
}

int hrt_retarget_body_quat(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters,
                           float damping, float rot_weight, float* d_robot_local_q, float* d_dof,
                           float* d_link_pos, void* stream) {
    HRT_ENTER(ctx);
    BodyQuatArgs a;
    if ((rc = fill_body_quat_args(ctx, B, d_src_gq, flags, ik_iters, damping, rot_weight, d_robot_local_q, d_dof,
                                  d_link_pos, &a)))
        return rc;
    if (B == 0) return 0;
    if (!d_src_gq) return fail(HRT_E_INVALID_ARG, "null input");
    if (!aligned16(d_src_gq) || !aligned16(d_robot_local_q) || !aligned16(d_dof) || !aligned16(d_link_pos))
        return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    return launch_body_quat(ctx, a, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU reassembly over NVLink, one process per GPU (BASELINE configs[4]): plain cudaMalloc buffers shared by
// CUDA IPC, filled by the compute kernel's own TMA bulk stores; a flag exchange closes the step.
// ---------------------------------------------------------------------------------------------
int hrt_peer_alloc(hrt_ctx* ctx, size_t bytes, void** d_ptr, unsigned char* handle64) {
    HRT_ENTER(ctx);
    if (!d_ptr || !handle64 || bytes == 0) return fail(HRT_E_INVALID_ARG, "null pointer / zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == HRT_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    HRT_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail((int)e, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle64, &h, sizeof(h));
    *d_ptr = p;
    return 0;
}

int hrt_peer_free(hrt_ctx* ctx, void* d_ptr) {
    HRT_ENTER(ctx);
    if (d_ptr) HRT_CUDA(cudaFree(d_ptr));
    return 0;
}

int hrt_peer_open(hrt_ctx* ctx, const unsigned char* handle64, void** d_ptr) {
    HRT_ENTER(ctx);
    if (!d_ptr || !handle64) return fail(HRT_E_INVALID_ARG, "null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    HRT_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int hrt_peer_close(hrt_ctx* ctx, void* d_ptr) {
    HRT_ENTER(ctx);
    if (d_ptr) HRT_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int hrt_retarget_body_quat_gather(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters, float damping,
                                  float rot_weight, float* d_link_pos, int n_peer, float* const* d_peer_dof, int64_t frame0,
                                  void* stream) {
    HRT_ENTER(ctx);
    BodyQuatArgs a;
    if ((rc = fill_body_quat_args(ctx, B, d_src_gq, flags, ik_iters, damping, rot_weight, nullptr, nullptr, d_link_pos, &a))) return rc;
    if (n_peer < 1 || n_peer > HRT_MAX_PEERS || !d_peer_dof) return fail(HRT_E_INVALID_ARG, "n_peer must be 1..%d", HRT_MAX_PEERS);
    if (B == 0) return 0;
    if (frame0 < 0 || (frame0 * (ctx->bq.J_rob - 1) * 4) % 16 != 0)
        return fail(HRT_E_ALIGNMENT, "frame0 must keep the dof rows 16-byte aligned (a multiple of 4 frames for 30 DOFs)");
    if (!d_src_gq) return fail(HRT_E_INVALID_ARG, "null input");
    if (!aligned16(d_src_gq) || !aligned16(d_link_pos)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    a.flags &= ~BQ_PACKED_IK;
    a.n_peer = n_peer;
    a.peer_frame0 = frame0;
    for (int r = 0; r < n_peer; ++r) {
        if (!d_peer_dof[r] || !aligned16(d_peer_dof[r])) return fail(HRT_E_ALIGNMENT, "peer buffer %d null or not 16-byte aligned", r);
        a.peer_dof[r] = d_peer_dof[r];
    }
    return launch_body_quat(ctx, a, (cudaStream_t)stream);
}

int hrt_retarget_body_quat_multicast(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters, float damping,
                                     float rot_weight, float* d_link_pos, float* d_mc_dof, int64_t frame0, void* stream) {
    HRT_ENTER(ctx);
    BodyQuatArgs a;
    if ((rc = fill_body_quat_args(ctx, B, d_src_gq, flags, ik_iters, damping, rot_weight, nullptr, nullptr, d_link_pos, &a))) return rc;
    if (!d_mc_dof || !aligned16(d_mc_dof)) return fail(HRT_E_ALIGNMENT, "multicast address null or not 16-byte aligned");
    if (B == 0) return 0;
    if (frame0 < 0 || (frame0 * (ctx->bq.J_rob - 1) * 4) % 16 != 0)
        return fail(HRT_E_ALIGNMENT, "frame0 must keep the dof rows 16-byte aligned (a multiple of 4 frames for 30 DOFs)");
    if (!d_src_gq) return fail(HRT_E_INVALID_ARG, "null input");
    if (!aligned16(d_src_gq) || !aligned16(d_link_pos)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    a.flags &= ~BQ_PACKED_IK;
    a.peer_frame0 = frame0;
    a.mc_dof = d_mc_dof;
    return launch_body_quat(ctx, a, (cudaStream_t)stream);
}

int hrt_reassembly_layout(hrt_ctx* ctx, int64_t n_total, int n_rank, const int64_t* shard_frames, size_t* staging_bytes,
                          size_t* flag_offset, size_t* total_bytes, int* max_rounds) {
    HRT_ENTER(ctx);
    if (n_total < 0 || n_rank < 1 || n_rank > HRT_MAX_PEERS || !shard_frames) return fail(HRT_E_INVALID_ARG, "bad shard description");
    const long long T = (long long)ctx->sm_count * BQ_WARPS_WIDE;
    long long rounds = 1;
    for (int r = 0; r < n_rank; ++r) {
        const long long g = (shard_frames[r] + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
        rounds = std::max(rounds, (g + T - 1) / T);
    }
    if (rounds > 0x3fffffff) return fail(HRT_E_INVALID_ARG, "shard too long for the in-kernel reassembly");
    // one self-validating group (16 frames x 14 hinge angles + a check block) per 16 clip frames; no flags
    const size_t groups = (size_t)((std::max<int64_t>(n_total, 1) + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP);
    size_t stage = groups * BQ_PK_GROUP * sizeof(float);
    stage = (stage + 255) / 256 * 256;
    if (staging_bytes) *staging_bytes = stage;
    if (flag_offset) *flag_offset = stage;
    if (total_bytes) *total_bytes = stage;
    if (max_rounds) *max_rounds = (int)rounds;
    return 0;
}

int hrt_retarget_body_quat_reassemble(hrt_ctx* ctx, int64_t B, const float* d_src_gq, unsigned flags, int ik_iters, float damping,
                                      float rot_weight, float* d_link_pos, float* d_full_dof, int64_t n_total, int n_rank, int my_rank,
                                      const int64_t* shard_lo, const int64_t* shard_frames, void* d_symm, void* d_symm_mc,
                                      unsigned epoch, void* stream) {
    HRT_ENTER(ctx);
    if (n_rank < 1 || n_rank > HRT_MAX_PEERS || my_rank < 0 || my_rank >= n_rank || !shard_lo || !shard_frames)
        return fail(HRT_E_INVALID_ARG, "bad rank / shard description");
    if (!d_full_dof || !d_symm || !d_symm_mc) return fail(HRT_E_INVALID_ARG, "null reassembly buffer");
    if (B != shard_frames[my_rank]) return fail(HRT_E_INVALID_ARG, "B must be this rank's shard length");
    if (ctx->bq.arm[0].rob_first + 7 > ctx->bq.arm[1].rob_first && ctx->bq.arm[1].rob_first + 7 > ctx->bq.arm[0].rob_first)
        return fail(HRT_E_INVALID_ARG, "the two arms' hinge ranges overlap");
    if ((ctx->bq.J_rob - 1) % 2) return fail(HRT_E_INVALID_ARG, "an odd DOF count is not supported by the in-kernel reassembly");
    BodyQuatArgs a;
    // B == 0 is a legal shard here: the rank still unpacks what its peers send
    if ((rc = fill_body_quat_args(ctx, std::max<int64_t>(B, 0), d_src_gq, flags, ik_iters, damping, rot_weight, nullptr,
                                  d_full_dof + shard_lo[my_rank] * (ctx->bq.J_rob - 1), d_link_pos, &a))) return rc;
    if (B > 0 && !d_src_gq) return fail(HRT_E_INVALID_ARG, "null input");
    if (!aligned16(d_src_gq) || !aligned16(d_link_pos) || !aligned16(d_full_dof) || !aligned16(d_symm) || !aligned16(d_symm_mc))
        return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    a.flags &= ~BQ_PACKED_IK;
    if ((rc = hrt_reassembly_layout(ctx, n_total, n_rank, shard_frames, nullptr, nullptr, nullptr, nullptr))) return rc;
    if (epoch == 0u) return fail(HRT_E_INVALID_ARG, "epoch 0 is the zero-filled staging buffers' own: start at 1");
    int64_t covered = 0;
    for (int r = 0; r < n_rank; ++r) {
        if ((shard_frames[r] > 0 && shard_lo[r] % BQ_FRAMES_PER_WARP) || shard_lo[r] != covered || shard_frames[r] < 0)
            return fail(HRT_E_INVALID_ARG, "shards must tile the clip in rank order and start at multiples of %d frames", BQ_FRAMES_PER_WARP);
        covered += shard_frames[r];
        a.g.lo[r] = shard_lo[r];
        a.g.n[r] = shard_frames[r];
    }
    if (covered != n_total) return fail(HRT_E_INVALID_ARG, "shards do not add up to the clip");
    a.g.mc_pk = (float*)d_symm_mc;
    a.g.pk = (const float*)d_symm;
    a.g.full = d_full_dof;
    a.g.n_rank = n_rank;
    a.g.me = my_rank;
    // equal shard strides (what sharding.shard_range produces) let the kernel derive every peer address from a group index
    a.g.shard_groups = 0;
    if (n_rank > 1 && shard_lo[1] > 0 && shard_lo[1] % BQ_FRAMES_PER_WARP == 0 && n_total / BQ_FRAMES_PER_WARP < (1LL << 30)) {
        bool affine = true;
        for (int r = 0; r < n_rank; ++r) affine = affine && shard_lo[r] == (int64_t)r * shard_lo[1];
        if (affine) a.g.shard_groups = (int)(shard_lo[1] / BQ_FRAMES_PER_WARP);
    }
    a.g.epoch = epoch;
    a.g.timeout_ns = 20ull * 1000 * 1000 * 1000;
    {
        static const unsigned dbg = [] { const char* e = getenv("HRT_GATHER_DEBUG"); return e ? (unsigned)atoi(e) : 0u; }();
        a.g.debug = dbg;
        // no multicast mapping (d_symm_mc == d_symm, which a real multicast address never is): nothing is published; the
        // rank only unpacks what is already in its staging buffer (single-process replay of a rank, used by the tests)
        if (d_symm_mc == d_symm) a.g.debug |= 1u;
    }
    // every rank launches the same geometry (one CTA per SM, 16 warps): a group's round and warp follow from its index
    const size_t smem = ((size_t)BQ_CONST_WORDS + (size_t)BQ_WARPS_WIDE * bq_tile_words(ctx->bq.J_src, ctx->bq.J_rob, false) +
                         (size_t)bq_gather_words(BQ_WARPS_WIDE)) * sizeof(float);
    auto kern = body_quat_gather_kernel<BQ_WARPS_WIDE>;
    static size_t attr_done[16] = {0};
    const int d = ctx->device & 15;
    if (attr_done[d] < smem) {
        HRT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done[d] = smem;
    }
    kern<<<ctx->sm_count, (BQ_WARPS_WIDE + BQ_GATHER_UNPACK_WARPS) * 32, smem, (cudaStream_t)stream>>>(ctx->bq, a);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_peer_barrier(hrt_ctx* ctx, int n_peer, int my_rank, unsigned* const* d_peer_flags, unsigned epoch, void* stream) {
    HRT_ENTER(ctx);
    if (n_peer < 1 || n_peer > HRT_MAX_PEERS || my_rank < 0 || my_rank >= n_peer || !d_peer_flags)
        return fail(HRT_E_INVALID_ARG, "bad peer set");
    PeerFlags pf;
    for (int r = 0; r < HRT_MAX_PEERS; ++r) pf.flags[r] = r < n_peer ? d_peer_flags[r] : nullptr;
    for (int r = 0; r < n_peer; ++r)
        if (!pf.flags[r]) return fail(HRT_E_INVALID_ARG, "peer flag array %d is null", r);
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, n_peer, my_rank, epoch, 20ull * 1000 * 1000 * 1000);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_retarget_body_quat_host(hrt_ctx* ctx, int64_t B, const float* h_src_gq, unsigned flags, int ik_iters,
                                float damping, float rot_weight, float* h_robot_local_q, float* h_dof,
                                float* h_link_pos) {
    HRT_ENTER(ctx);
    BodyQuatArgs proto;
    if ((rc = fill_body_quat_args(ctx, B, nullptr, flags, ik_iters, damping, rot_weight, nullptr, nullptr, nullptr, &proto)))
        return rc;
    if (B == 0) return 0;
    if (!h_src_gq) return fail(HRT_E_INVALID_ARG, "null input");
    const int JS = ctx->bq.J_src, JR = ctx->bq.J_rob;
    const size_t in_b = (size_t)JS * 16, lq_b = (size_t)JR * 16, dof_b = (size_t)(JR - 1) * 4, lp_b = (size_t)JR * 12;
    // frames per pipeline stage (multiple of 16); HRT_HOST_CHUNK_LOG2 overrides it for tuning runs
    static const long long chunk = [] {
        const char* e = std::getenv("HRT_HOST_CHUNK_LOG2");
        const int lg = e ? std::atoi(e) : 16;
        return 1LL << (lg < 10 ? 10 : (lg > 20 ? 20 : lg));
    }();
    const size_t per_frame = in_b + lq_b + dof_b + lp_b;
    const size_t need = per_frame * chunk;
    // The solver reads 9 of the source joints (torso + 4 per arm).  When the range of joints that holds them is well short
    // of a row (vtrdyn: joints 10-20, 176 of 336 bytes), only that column range crosses PCIe, as one strided copy per
    // chunk: the copy engine moves ~210 M rows/s (measured, tools/memcpy2d_probe.py: 2^20 rows of 176 B in 4.9 ms against
    // 7.1 ms for the whole rows).  The device rows' other columns are never read by the kernel (zeroed once, below).
    int jmin = JS, jmax = -1;
    for (int side = 0; side < 2; ++side) {
        const ArmParams& ap = ctx->bq.arm[side];
        for (int j : {ap.src_torso, ap.src_shoulder, ap.src_upper, ap.src_lower, ap.src_hand}) { jmin = std::min(jmin, j); jmax = std::max(jmax, j); }
    }
    const size_t col_off = (size_t)jmin * 16, col_w = (size_t)(jmax - jmin + 1) * 16;
    // Taken when the inputs are the larger side of the traffic (dof-only calls: 2^20 frames 1.54e8 -> 1.84e8 frames/s): with the
    // link positions or local rotations going back the call is bound by the device-to-host direction, which the strided
    // reads slow down (9.7e7 -> 8.6e7 frames/s measured), so whole rows are copied then.
    const size_t out_b = (h_robot_local_q ? lq_b : 0) + (h_dof ? dof_b : 0) + (h_link_pos ? lp_b : 0);
    const bool columns = col_w * 4 <= in_b * 3 && col_w >= 128 && out_b < in_b;            // narrower rows run into the per-row cost
    if (ctx->d_stage_bytes < need) {
        for (int i = 0; i < kHostStreams; ++i) {
            if (ctx->d_stage[i]) { cudaFree(ctx->d_stage[i]); ctx->d_stage[i] = nullptr; }
            HRT_CUDA(cudaMalloc(&ctx->d_stage[i], need));
            HRT_CUDA(cudaMemset(ctx->d_stage[i], 0, need));
            if (!ctx->hs[i]) HRT_CUDA(cudaStreamCreateWithFlags(&ctx->hs[i], cudaStreamNonBlocking));
            if (!ctx->hs_done[i]) HRT_CUDA(cudaEventCreateWithFlags(&ctx->hs_done[i], cudaEventDisableTiming));
        }
        ctx->d_stage_bytes = need;
    }
    int slot = 0;
    for (long long f0 = 0; f0 < B; f0 += chunk, slot = (slot + 1) % kHostStreams) {
        const long long n = std::min(chunk, (long long)B - f0);
        cudaStream_t st = ctx->hs[slot];
        char* base = reinterpret_cast<char*>(ctx->d_stage[slot]);
        float* d_in = reinterpret_cast<float*>(base);
        float* d_lq = reinterpret_cast<float*>(base + in_b * chunk);
        float* d_dof = reinterpret_cast<float*>(base + (in_b + lq_b) * chunk);
        float* d_lp = reinterpret_cast<float*>(base + (in_b + lq_b + dof_b) * chunk);
        // stream order on `st` already serialises reuse of this staging slot
        if (columns)
            HRT_CUDA(cudaMemcpy2DAsync(reinterpret_cast<char*>(d_in) + col_off, in_b, reinterpret_cast<const char*>(h_src_gq) + f0 * in_b + col_off,
                                       in_b, col_w, (size_t)n, cudaMemcpyHostToDevice, st));
        else
            HRT_CUDA(cudaMemcpyAsync(d_in, reinterpret_cast<const char*>(h_src_gq) + f0 * in_b, n * in_b, cudaMemcpyHostToDevice, st));
        BodyQuatArgs a = proto;
        a.B = n;
        a.src_gq = d_in;
        a.out_local_q = h_robot_local_q ? d_lq : nullptr;
        a.out_dof = h_dof ? d_dof : nullptr;
        a.out_link_pos = h_link_pos ? d_lp : nullptr;
        if ((rc = launch_body_quat(ctx, a, st))) return rc;
        if (h_robot_local_q)
            HRT_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(h_robot_local_q) + f0 * lq_b, d_lq, n * lq_b, cudaMemcpyDeviceToHost, st));
        if (h_dof)
            HRT_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(h_dof) + f0 * dof_b, d_dof, n * dof_b, cudaMemcpyDeviceToHost, st));
        if (h_link_pos)
            HRT_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(h_link_pos) + f0 * lp_b, d_lp, n * lp_b, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kHostStreams; ++i) HRT_CUDA(cudaStreamSynchronize(ctx->hs[i]));
    return 0;
}

int hrt_configure_pos(hrt_ctx* ctx, int mode, int src_tree, int rob_tree, const float* src_global_t, int precise_gripper) {
    HRT_ENTER(ctx);
    hrt_stream_pos_close(ctx);         // a stream (or resident server) opened before keeps the old PosParams by value
    if (mode < 0 || mode > 3) return fail(HRT_E_INVALID_ARG, "mode must be 0 (full_body_pos), 1 (upper_body), 2 (full_body) or 3 (main)");
    Tree *s, *r;
    if ((rc = get_tree(ctx, src_tree, &s))) return rc;
    if ((rc = get_tree(ctx, rob_tree, &r))) return rc;
    const TreeParams& st = s->tp;
    PosParams& pp = ctx->pos[mode];
    memset(&pp, 0, sizeof(pp));
    bool pos_limits_ok = false;
    pp.mode = mode;
    pp.J_rob = r->tp.J;
    if (pp.J_rob != 31) return fail(HRT_E_UNSUPPORTED_TREE, "the position-path solvers write Hu v5 joints 12-29 (31-joint robot)");
    pp.n_body = 21;
    pp.n_bodyq = 21;
    pp.n_hand = 20;
    pp.precise_gripper = precise_gripper ? 1 : 0;
    auto off = [&](int j, float* dst) { for (int k = 0; k < 3; ++k) dst[k] = st.jr[j].off[k]; };
    // body indices are the vtrdyn 21-joint order in every mode (retarget_solver.py:49-86,
    // full_body_pos_retargeter.py:68-110, full_body_retargeter.py:59-99)
    const int tp3[3] = {17, 13, 11};
    for (int n = 0; n < 3; ++n) pp.torso_pts[n] = tp3[n];
    pp.torso_org = 10;
    pp.bq_torso = 10;
    const int hk[5] = {2, 6, 10, 14, 17};
    for (int n = 0; n < 5; ++n) pp.hand_kabsch[n] = hk[n];
    pp.hand_org = 0;
    pp.flip[0] = pp.flip[1] = pp.flip[2] = 1.f;
    const int body_arm[2][3] = {{18, 19, 20}, {14, 15, 16}};
    const int rob_first[2] = {12, 21};
    for (int side = 0; side < 2; ++side) {
        PosArm& ar = pp.arm[side];
        ar.b_sh = body_arm[side][0]; ar.b_el = body_arm[side][1]; ar.b_wr = body_arm[side][2];
        ar.rob_first = rob_first[side];
        ar.q_parent = side == 0 ? 17 : 13;
        ar.q_wrist = side == 0 ? 20 : 16;
        ar.bq_wrist = side == 0 ? 14 : 39;
    }
    if (mode == POS_UPPER_BODY || mode == POS_MAIN) {
        if (st.J != 21) return fail(HRT_E_UNSUPPORTED_TREE, "upper-body / main solvers need the 21-joint vtrdyn zero pose");
        const int zt[3] = {17, 13, 11};                                   // retarget_solver.py:50
        for (int n = 0; n < 3; ++n) off(zt[n], pp.ztorso[n]);
        off(19, pp.arm[0].v0_upper); off(20, pp.arm[0].v0_lower);         // :56,76
        off(15, pp.arm[1].v0_upper); off(16, pp.arm[1].v0_lower);         // :62,84
        if (mode == POS_UPPER_BODY) { pp.flip[0] = -1.f; pp.flip[1] = -1.f; }   // :41 (main.py:170 flips before the rescale)
        if (mode == POS_MAIN) pp.arm[0].q_parent = pp.arm[1].q_parent = 10;     // main.py:206,212
    } else {
        if (st.J != 59) return fail(HRT_E_UNSUPPORTED_TREE, "full-body solvers need the 59-joint vtrdyn_full zero pose");
        const int zt[3] = {11, 36, 34};                                   // full_body_pos_retargeter.py:69
        for (int n = 0; n < 3; ++n) off(zt[n], pp.ztorso[n]);
        off(13, pp.arm[0].v0_upper); off(14, pp.arm[0].v0_lower);         // :78,86
        off(38, pp.arm[1].v0_upper); off(39, pp.arm[1].v0_lower);         // :99,107
        const int zl[5] = {16, 20, 24, 28, 32}, zr[5] = {41, 45, 49, 53, 56};   // :139,162
        for (int n = 0; n < 5; ++n) { off(zl[n], pp.arm[0].zwrist[n]); off(zr[n], pp.arm[1].zwrist[n]); }
        const int tips_pos[5] = {4, 8, 12, 16, 19}, tips_full[5] = {3, 7, 11, 15, 19};
        const int ext[5] = {18, 22, 26, 30, 33};
        float sum = 0.f;
        if (mode == POS_FULL_BODY_POS) {
            if (!src_global_t) return fail(HRT_E_INVALID_ARG, "full_body_pos needs the source zero pose's global translations");
            pp.J_bq = 59;
            for (int n = 0; n < 5; ++n) {
                pp.hand_tips[n] = tips_pos[n];
                const float d = src_global_t[ext[n] * 3] - src_global_t[14 * 3];        // :184
                sum = n == 0 ? d : sum + d;
            }
        } else {
            for (int n = 0; n < 5; ++n) {
                pp.hand_tips[n] = tips_full[n];
                const float d = st.jr[ext[n]].off[0] - st.jr[24].off[0];                // full_body_retargeter.py:152
                sum = n == 0 ? d : sum + d;
            }
        }
        pp.orig_x = sum / 5.f;
    }
    {
        // robot-side arm tables for the limit-aware refinement: zero-pose positions p_j = off_j + p_parent
        const TreeParams& rt = r->tp;
        pos_limits_ok = true;
        std::vector<float> pos(rt.J * 3, 0.f);
        for (int j = 1; j < rt.J; ++j)
            for (int k = 0; k < 3; ++k) pos[j * 3 + k] = rt.jr[j].off[k] + pos[rt.parent[j] * 3 + k];
        for (int side = 0; side < 2; ++side) {
            const int f = rob_first[side];
            for (int c = 0; c < 9; ++c)
                for (int k = 0; k < 3; ++k) pp.ik[side].off[c][k] = rt.jr[f + c].off[k];
            for (int c = 0; c < 7; ++c) {
                pp.ik[side].lower[c] = rt.lim[f + c][0];
                pp.ik[side].upper[c] = rt.lim[f + c][1];
                if (!(pp.ik[side].lower[c] >= -kArmLimitMax && pp.ik[side].upper[c] <= kArmLimitMax)) pos_limits_ok = false;
            }
            for (int k = 0; k < 3; ++k) pp.ik[side].p_sh[k] = pos[f * 3 + k];
        }
    }
    {
        // zero-pose bone angles: computed on the device once (the frames' own code path), kept in the parameter block
        float* d_za = nullptr;
        HRT_CUDA(cudaMalloc(&d_za, 8 * sizeof(float)));
        pos_zero_angles_kernel<<<1, 32>>>(pp.arm[0], pp.arm[1], d_za);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(&pp.zero_ang[0][0], d_za, 8 * sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(d_za);
        if (e != cudaSuccess) return fail((int)e, "zero-pose bone angles: %s", cudaGetErrorString(e));
    }
    ctx->pos_set[mode] = true;
    ctx->pos_limits_ok[mode] = pos_limits_ok;
    if (mode == POS_FULL_BODY_POS) {
        ctx->pos_limits_ok[4] = pos_limits_ok;
        // the same solver reading the mocap wire layout (sim_full_body_teleop.py:109-112): body rows
        // 23 -> 21 and the HandNodes -> solver finger order are index remaps, applied to the tables once
        static const int body_map[21] = {0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22};
        static const int hand_map[20] = {0, 4, 5, 6, 7, 8, 9, 10, 11, 16, 17, 18, 19, 12, 13, 14, 15, 1, 2, 3};
        PosParams& w = ctx->pos[4];
        w = pp;
        w.n_body = 23;
        for (int n = 0; n < 3; ++n) w.torso_pts[n] = body_map[pp.torso_pts[n]];
        w.torso_org = body_map[pp.torso_org];
        for (int n = 0; n < 5; ++n) { w.hand_kabsch[n] = hand_map[pp.hand_kabsch[n]]; w.hand_tips[n] = hand_map[pp.hand_tips[n]]; }
        w.hand_org = hand_map[pp.hand_org];
        for (int side = 0; side < 2; ++side) {
            w.arm[side].b_sh = body_map[pp.arm[side].b_sh];
            w.arm[side].b_el = body_map[pp.arm[side].b_el];
            w.arm[side].b_wr = body_map[pp.arm[side].b_wr];
        }
        ctx->pos_set[4] = true;
    }
    return 0;
}

static int launch_pos(hrt_ctx* ctx, int slot, const PosArgs& a, cudaStream_t st, int force_grid = 0) {
    if (!ctx->pos_set[slot]) return fail(HRT_E_NOT_CONFIGURED, "hrt_configure_pos(mode %d) has not been called", slot == 4 ? 0 : slot);
    const int mode = ctx->pos[slot].mode;
    if ((a.flags & (POS_CLAMP | POS_IK)) && !ctx->pos_limits_ok[slot])
        return fail(HRT_E_UNSUPPORTED_TREE, "limits / refinement need finite arm hinge limits that do not reach beyond +-pi in the robot tree");
    if (a.B < 0) return fail(HRT_E_INVALID_ARG, "negative frame count");
    if (a.B == 0) return 0;
    const void* ptrs[] = {a.body_t, a.lhand_t, a.rhand_t, a.body_q, a.out_local_q, a.out_dof, a.out_body_gq};
    for (const void* p : ptrs)
        if (!aligned16(p)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    const PosParams& pp = ctx->pos[slot];
    const bool with_bq = mode == POS_FULL_BODY_POS && a.out_body_gq;
    const size_t tile_bytes = (size_t)pos_tile_words(pp, a.out_local_q != nullptr, with_bq) * sizeof(float);
    const size_t const_bytes = (size_t)pos_const_words() * sizeof(float);
    const long long groups = (a.B + BQ_FRAMES_PER_WARP - 1) / BQ_FRAMES_PER_WARP;
    // the most warps per CTA whose staging tiles fit (a latency-bound kernel: see POS_WARPS_* in hrt_pos.cuh); a call
    // that cannot fill one small CTA per SM (single frames, short clips) stays with the small CTA
    int warps = POS_WARPS_MIN;
    if (groups > (long long)ctx->sm_count * POS_WARPS_MIN) {
        if (const_bytes + POS_WARPS_MAX * tile_bytes <= 226 * 1024) warps = POS_WARPS_MAX;
        else if (const_bytes + POS_WARPS_MID * tile_bytes <= 226 * 1024) warps = POS_WARPS_MID;
    }
    const size_t smem = const_bytes + (size_t)warps * tile_bytes;
    const long long ctas = (groups + warps - 1) / warps;
    const int grid = force_grid ? force_grid : (int)std::max(1LL, std::min(ctas, (long long)ctx->sm_count));
#define HRT_POS_LAUNCH(MODE)                                                                                             \
    do {                                                                                                                 \
        if (warps == POS_WARPS_MAX) pos_retarget_kernel<MODE, POS_WARPS_MAX><<<grid, POS_WARPS_MAX * 32, smem, st>>>(pp, a);      \
        else if (warps == POS_WARPS_MID) pos_retarget_kernel<MODE, POS_WARPS_MID><<<grid, POS_WARPS_MID * 32, smem, st>>>(pp, a); \
        else pos_retarget_kernel<MODE, POS_WARPS_MIN><<<grid, POS_WARPS_MIN * 32, smem, st>>>(pp, a);                             \
    } while (0)
    if (mode == POS_MAIN) HRT_POS_LAUNCH(POS_MAIN);
    else if (mode == POS_FULL_BODY_POS) HRT_POS_LAUNCH(POS_FULL_BODY_POS);
    else if (mode == POS_UPPER_BODY) HRT_POS_LAUNCH(POS_UPPER_BODY);
    else HRT_POS_LAUNCH(POS_FULL_BODY);
#undef HRT_POS_LAUNCH
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_retarget_full_body_pos(hrt_ctx* ctx, int64_t B, const float* d_body_t, const float* d_lhand_t, const float* d_rhand_t,
                               float* d_robot_local_q, float* d_dof, float* d_body_gq, void* stream) {
    HRT_ENTER(ctx);
    if (B > 0 && (!d_body_t || !d_lhand_t || !d_rhand_t)) return fail(HRT_E_INVALID_ARG, "null input");
    PosArgs a{};
    a.B = B; a.body_t = d_body_t; a.lhand_t = d_lhand_t; a.rhand_t = d_rhand_t;
    a.out_local_q = d_robot_local_q; a.out_dof = d_dof; a.out_body_gq = d_body_gq;
    return launch_pos(ctx, POS_FULL_BODY_POS, a, (cudaStream_t)stream);
}

int hrt_retarget_full_body_pos_ex(hrt_ctx* ctx, int64_t B, const float* d_body_t, const float* d_lhand_t, const float* d_rhand_t,
                                  unsigned flags, int ik_iters, float damping, float rot_weight, float* d_robot_local_q,
                                  float* d_dof, float* d_body_gq, void* stream) {
    HRT_ENTER(ctx);
    if (B > 0 && (!d_body_t || !d_lhand_t || !d_rhand_t)) return fail(HRT_E_INVALID_ARG, "null input");
    if ((flags & HRT_POS_IK) && (ik_iters < 0 || ik_iters > 1000)) return fail(HRT_E_INVALID_ARG, "ik_iters out of range");
    PosArgs a{};
    a.B = B; a.body_t = d_body_t; a.lhand_t = d_lhand_t; a.rhand_t = d_rhand_t;
    a.out_local_q = d_robot_local_q; a.out_dof = d_dof; a.out_body_gq = d_body_gq;
    a.flags = (flags & HRT_POS_CLAMP ? POS_CLAMP : 0u) | (flags & HRT_POS_IK ? POS_IK : 0u);
    a.ik_iters = ik_iters; a.damping = damping; a.rot_weight = rot_weight;
    return launch_pos(ctx, POS_FULL_BODY_POS, a, (cudaStream_t)stream);
}

int hrt_retarget_full_body_pos_host(hrt_ctx* ctx, int64_t B, const float* h_body_t, const float* h_lhand_t, const float* h_rhand_t,
                                    unsigned flags, int ik_iters, float damping, float rot_weight, float* h_robot_local_q,
                                    float* h_dof) {
    HRT_ENTER(ctx);
    if (!ctx->pos_set[POS_FULL_BODY_POS]) return fail(HRT_E_NOT_CONFIGURED, "hrt_configure_pos(mode 0) has not been called");
    if (B < 0) return fail(HRT_E_INVALID_ARG, "negative frame count");
    if ((flags & HRT_POS_IK) && (ik_iters < 0 || ik_iters > 1000)) return fail(HRT_E_INVALID_ARG, "ik_iters out of range");
    if (B == 0) return 0;
    if (!h_body_t || !h_lhand_t || !h_rhand_t) return fail(HRT_E_INVALID_ARG, "null input");
    const PosParams& pp = ctx->pos[POS_FULL_BODY_POS];
    const size_t body_b = (size_t)pp.n_body * 12, hand_b = (size_t)pp.n_hand * 12, lq_b = (size_t)pp.J_rob * 16, dof_b = (size_t)(pp.J_rob - 1) * 4;
    // frames per pipeline stage (multiple of 16); HRT_HOST_CHUNK_LOG2 overrides it for tuning runs
    static const long long chunk = [] {
        const char* e = std::getenv("HRT_HOST_CHUNK_LOG2");
        const int lg = e ? std::atoi(e) : 16;
        return 1LL << (lg < 10 ? 10 : (lg > 20 ? 20 : lg));
    }();
    const size_t need = (body_b + 2 * hand_b + lq_b + dof_b) * chunk;
    if (ctx->d_stage_bytes < need) {
        for (int i = 0; i < kHostStreams; ++i) {
            if (ctx->d_stage[i]) { cudaFree(ctx->d_stage[i]); ctx->d_stage[i] = nullptr; }
            HRT_CUDA(cudaMalloc(&ctx->d_stage[i], need));
            if (!ctx->hs[i]) HRT_CUDA(cudaStreamCreateWithFlags(&ctx->hs[i], cudaStreamNonBlocking));
            if (!ctx->hs_done[i]) HRT_CUDA(cudaEventCreateWithFlags(&ctx->hs_done[i], cudaEventDisableTiming));
        }
        ctx->d_stage_bytes = need;
    }
    int slot = 0;
    for (long long f0 = 0; f0 < B; f0 += chunk, slot = (slot + 1) % kHostStreams) {
        const long long n = std::min(chunk, (long long)B - f0);
        cudaStream_t st = ctx->hs[slot];
        char* base = reinterpret_cast<char*>(ctx->d_stage[slot]);
        float* d_body = reinterpret_cast<float*>(base);
        float* d_lh = reinterpret_cast<float*>(base + body_b * chunk);
        float* d_rh = reinterpret_cast<float*>(base + (body_b + hand_b) * chunk);
        float* d_lq = reinterpret_cast<float*>(base + (body_b + 2 * hand_b) * chunk);
        float* d_dof = reinterpret_cast<float*>(base + (body_b + 2 * hand_b + lq_b) * chunk);
        HRT_CUDA(cudaMemcpyAsync(d_body, reinterpret_cast<const char*>(h_body_t) + f0 * body_b, n * body_b, cudaMemcpyHostToDevice, st));
        HRT_CUDA(cudaMemcpyAsync(d_lh, reinterpret_cast<const char*>(h_lhand_t) + f0 * hand_b, n * hand_b, cudaMemcpyHostToDevice, st));
        HRT_CUDA(cudaMemcpyAsync(d_rh, reinterpret_cast<const char*>(h_rhand_t) + f0 * hand_b, n * hand_b, cudaMemcpyHostToDevice, st));
        PosArgs a{};
        a.B = n; a.body_t = d_body; a.lhand_t = d_lh; a.rhand_t = d_rh;
        a.out_local_q = h_robot_local_q ? d_lq : nullptr;
        a.out_dof = h_dof ? d_dof : nullptr;
        a.flags = (flags & HRT_POS_CLAMP ? POS_CLAMP : 0u) | (flags & HRT_POS_IK ? POS_IK : 0u);
        a.ik_iters = ik_iters; a.damping = damping; a.rot_weight = rot_weight;
        if ((rc = launch_pos(ctx, POS_FULL_BODY_POS, a, st))) return rc;
        if (h_robot_local_q)
            HRT_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(h_robot_local_q) + f0 * lq_b, d_lq, n * lq_b, cudaMemcpyDeviceToHost, st));
        if (h_dof)
            HRT_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(h_dof) + f0 * dof_b, d_dof, n * dof_b, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kHostStreams; ++i) HRT_CUDA(cudaStreamSynchronize(ctx->hs[i]));
    return 0;
}

int hrt_retarget_upper_body(hrt_ctx* ctx, int64_t B, const float* d_body_t, float* d_robot_local_q, float* d_dof, void* stream) {
    HRT_ENTER(ctx);
    if (B > 0 && !d_body_t) return fail(HRT_E_INVALID_ARG, "null input");
    PosArgs a{};
    a.B = B; a.body_t = d_body_t; a.out_local_q = d_robot_local_q; a.out_dof = d_dof;
    return launch_pos(ctx, POS_UPPER_BODY, a, (cudaStream_t)stream);
}

int hrt_retarget_full_body(hrt_ctx* ctx, int64_t B, const float* d_body_q, const float* d_body_t, const float* d_lhand_t,
                           const float* d_rhand_t, float* d_robot_local_q, float* d_dof, void* stream) {
    HRT_ENTER(ctx);
    if (B > 0 && (!d_body_q || !d_body_t || !d_lhand_t || !d_rhand_t)) return fail(HRT_E_INVALID_ARG, "null input");
    PosArgs a{};
    a.B = B; a.body_q = d_body_q; a.body_t = d_body_t; a.lhand_t = d_lhand_t; a.rhand_t = d_rhand_t;
    a.out_local_q = d_robot_local_q; a.out_dof = d_dof;
    return launch_pos(ctx, POS_FULL_BODY, a, (cudaStream_t)stream);
}

int hrt_stream_open(hrt_ctx* ctx, unsigned flags, int ik_iters, float damping, float rot_weight) {
    HRT_ENTER(ctx);
    if (ctx->stream_open) hrt_stream_close(ctx);
    BodyQuatArgs a;
    if ((rc = fill_body_quat_args(ctx, 1, nullptr, flags, ik_iters, damping, rot_weight, nullptr, nullptr, nullptr, &a)))
        return rc;
    a.flags &= ~BQ_PACKED_IK;
    const int JS = ctx->bq.J_src, JR = ctx->bq.J_rob;
    const size_t in_w = (size_t)JS * 4;
    const size_t out_w = (size_t)JR * 4 + 32 + (size_t)JR * 3 + 3;   // local_q | dof (padded to 32) | link pos
    HRT_CUDA(cudaHostAlloc(&ctx->mb_in, in_w * sizeof(float), cudaHostAllocMapped));
    HRT_CUDA(cudaHostAlloc(&ctx->mb_out, out_w * sizeof(float), cudaHostAllocMapped));
    HRT_CUDA(cudaHostGetDevicePointer(&ctx->mb_in_d, ctx->mb_in, 0));
    HRT_CUDA(cudaHostGetDevicePointer(&ctx->mb_out_d, ctx->mb_out, 0));
    HRT_CUDA(cudaStreamCreateWithFlags(&ctx->ss, cudaStreamNonBlocking));
    a.src_gq = ctx->mb_in_d;
    a.out_local_q = ctx->mb_out_d;
    a.out_dof = ctx->mb_out_d + JR * 4;
    a.out_link_pos = ctx->mb_out_d + JR * 4 + 32;
    ctx->stream_args = a;
    ctx->stream_persistent = (flags & HRT_BQ_PERSISTENT) != 0;
    ctx->bserver_launched = false;
    ctx->bseq = 0;
    if (ctx->stream_persistent) {
        unsigned* c = nullptr;
        HRT_CUDA(cudaHostAlloc(&c, 32 * sizeof(unsigned), cudaHostAllocMapped));
        memset(c, 0, 32 * sizeof(unsigned));
        ctx->bctrl = c;
        HRT_CUDA(cudaHostGetDevicePointer(&ctx->bctrl_d, c, 0));
        HRT_CUDA(cudaFuncSetAttribute(bq_stream_server_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    ctx->stream_open = true;
    return 0;
}

int hrt_stream_frame(hrt_ctx* ctx, const float* h_src_gq, float* h_robot_local_q, float* h_dof,
                     float* h_link_pos) {
    if (!ctx || !ctx->stream_open) return fail(HRT_E_NOT_CONFIGURED, "hrt_stream_open has not been called");
    if (!h_src_gq) return fail(HRT_E_INVALID_ARG, "null input");
    DeviceGuard dev_guard_;                      // the launch on ctx->ss needs the context's device current
    if (!ctx->stream_persistent) {
        int rc = dev_guard_.enter(ctx);
        if (rc) return rc;
    }
    const int JS = ctx->bq.J_src, JR = ctx->bq.J_rob;
    memcpy(ctx->mb_in, h_src_gq, (size_t)JS * 16);
    if (!ctx->stream_persistent) {
        int rc = launch_body_quat(ctx, ctx->stream_args, ctx->ss, 1);
        if (rc) return rc;
        HRT_CUDA(cudaStreamSynchronize(ctx->ss));
    } else {
        const size_t smem = ((size_t)BQ_CONST_WORDS + bq_tile_words(JS, JR, true)) * sizeof(float);
        int rc = server_roundtrip(ctx->bctrl, &ctx->bseq, &ctx->bserver_launched, ctx->ss, [&](unsigned served) {
            bq_stream_server_kernel<<<1, 32, smem, ctx->ss>>>(ctx->bq, ctx->stream_args, ctx->bctrl_d, served, kServerIdleNs);
            HRT_CUDA(cudaGetLastError());
            return 0;
        });
        if (rc) return rc;
    }
    if (h_robot_local_q) memcpy(h_robot_local_q, ctx->mb_out, (size_t)JR * 16);
    if (h_dof) memcpy(h_dof, ctx->mb_out + JR * 4, (size_t)(JR - 1) * 4);
    if (h_link_pos) memcpy(h_link_pos, ctx->mb_out + JR * 4 + 32, (size_t)JR * 12);
    return 0;
}

int hrt_stream_close(hrt_ctx* ctx) {
    if (!ctx || !ctx->stream_open) return 0;
    DeviceGuard dev_guard_;
    dev_guard_.enter(ctx);
    if (ctx->bctrl) ctx->bctrl[1] = 1u;               // tell a resident server to leave
    if (ctx->ss) { cudaStreamSynchronize(ctx->ss); cudaStreamDestroy(ctx->ss); ctx->ss = nullptr; }
    if (ctx->mb_in) { cudaFreeHost(ctx->mb_in); ctx->mb_in = nullptr; }
    if (ctx->mb_out) { cudaFreeHost(ctx->mb_out); ctx->mb_out = nullptr; }
    if (ctx->bctrl) { cudaFreeHost(const_cast<unsigned*>(ctx->bctrl)); ctx->bctrl = nullptr; ctx->bctrl_d = nullptr; }
    ctx->stream_open = false;
    ctx->bserver_launched = false;
    return 0;
}

int hrt_retarget_full_body_pos_wire(hrt_ctx* ctx, int64_t B, const float* d_body23_t, const float* d_lhand_t, const float* d_rhand_t,
                                    float* d_robot_local_q, float* d_dof, float* d_body_gq, void* stream) {
    HRT_ENTER(ctx);
    if (B > 0 && (!d_body23_t || !d_lhand_t || !d_rhand_t)) return fail(HRT_E_INVALID_ARG, "null input");
    PosArgs a{};
    a.B = B; a.body_t = d_body23_t; a.lhand_t = d_lhand_t; a.rhand_t = d_rhand_t;
    a.out_local_q = d_robot_local_q; a.out_dof = d_dof; a.out_body_gq = d_body_gq;
    return launch_pos(ctx, 4, a, (cudaStream_t)stream);
}

int hrt_retarget_main_arms(hrt_ctx* ctx, int64_t B, const float* d_body_q, const float* d_body_t,
                           float* d_robot_local_q, float* d_dof, void* stream) {
    HRT_ENTER(ctx);
    if (B > 0 && (!d_body_q || !d_body_t)) return fail(HRT_E_INVALID_ARG, "null input");
    PosArgs a{};
    a.B = B; a.body_q = d_body_q; a.body_t = d_body_t; a.out_local_q = d_robot_local_q; a.out_dof = d_dof;
    return launch_pos(ctx, POS_MAIN, a, (cudaStream_t)stream);
}

namespace {

size_t pos_smem_bytes(const PosParams& pp, const PosArgs& a) {
    const bool with_bq = pp.mode == POS_FULL_BODY_POS && a.out_body_gq;
    return ((size_t)pos_const_words() + (size_t)pos_tile_words(pp, a.out_local_q != nullptr, with_bq)) * sizeof(float);   // one warp
}

int launch_pos_server(hrt_ctx* ctx, unsigned served) {
    const PosParams& pp = ctx->pos[ctx->pstream_mode];
    pos_stream_server_kernel<POS_FULL_BODY_POS><<<1, 32, pos_smem_bytes(pp, ctx->pstream_args), ctx->pss>>>(
        pp, ctx->pstream_args, ctx->pctrl_d, served, kServerIdleNs);
    HRT_CUDA(cudaGetLastError());
    return 0;
}
}  // namespace

int hrt_stream_pos_open(hrt_ctx* ctx, int flags) {
    HRT_ENTER(ctx);
    const int mode = (flags >> HRT_STREAM_MODE_SHIFT) & 3;
    const bool wire = (flags & HRT_STREAM_WIRE_LAYOUT) != 0;
    if (mode == POS_MAIN) return fail(HRT_E_INVALID_ARG, "streaming serves modes 0 (full_body_pos), 1 (upper_body), 2 (full_body)");
    if (mode != POS_FULL_BODY_POS && (flags & (HRT_STREAM_WIRE_LAYOUT | HRT_STREAM_PERSISTENT | HRT_STREAM_BODY_GQ)))
        return fail(HRT_E_INVALID_ARG, "wire layout, resident server and body quaternions exist for mode 0 only");
    const int slot = wire ? 4 : mode;
    if (!ctx->pos_set[slot]) return fail(HRT_E_NOT_CONFIGURED, "hrt_configure_pos(mode %d) has not been called", mode);
    if (ctx->pstream_open) hrt_stream_pos_close(ctx);
    const PosParams& pp = ctx->pos[slot];
    const bool hands = mode != POS_UPPER_BODY, quats = mode == POS_FULL_BODY;
    // mailbox in: body | lhand | rhand | body_q (16-byte aligned parts, a part a mode does not read stays unused)
    const size_t body_w = (size_t)(pp.n_body * 3 + 3) / 4 * 4, hand_w = (size_t)(pp.n_hand * 3 + 3) / 4 * 4;
    const size_t in_w = body_w + 2 * hand_w + (size_t)pp.n_bodyq * 4;
    const bool want_bq = (flags & HRT_STREAM_BODY_GQ) != 0;
    const size_t out_w = (size_t)pp.J_rob * 4 + 32 + (want_bq ? (size_t)pp.J_bq * 4 : 0);   // local_q | dof (padded) | body_gq
    HRT_CUDA(cudaHostAlloc(&ctx->pmb_in, in_w * sizeof(float), cudaHostAllocMapped));
    HRT_CUDA(cudaHostAlloc(&ctx->pmb_out, out_w * sizeof(float), cudaHostAllocMapped));
    HRT_CUDA(cudaHostGetDevicePointer(&ctx->pmb_in_d, ctx->pmb_in, 0));
    HRT_CUDA(cudaHostGetDevicePointer(&ctx->pmb_out_d, ctx->pmb_out, 0));
    HRT_CUDA(cudaStreamCreateWithFlags(&ctx->pss, cudaStreamNonBlocking));
    PosArgs a{};
    a.B = 1;
    a.body_t = ctx->pmb_in_d;
    if (hands) {
        a.lhand_t = ctx->pmb_in_d + body_w;
        a.rhand_t = a.lhand_t + hand_w;
    }
    if (quats) a.body_q = ctx->pmb_in_d + body_w + 2 * hand_w;
    a.out_local_q = ctx->pmb_out_d;
    a.out_dof = ctx->pmb_out_d + pp.J_rob * 4;
    a.out_body_gq = want_bq ? ctx->pmb_out_d + pp.J_rob * 4 + 32 : nullptr;
    a.flags = ((flags & HRT_STREAM_CLAMP) ? POS_CLAMP : 0u) | ((flags & HRT_STREAM_IK) ? (POS_CLAMP | POS_IK) : 0u);
    a.ik_iters = 10; a.damping = 0.1f; a.rot_weight = 0.2f;
    ctx->pstream_args = a;
    ctx->pstream_mode = slot;
    ctx->pstream_persistent = (flags & HRT_STREAM_PERSISTENT) != 0;
    ctx->pserver_launched = false;
    ctx->pseq = 0;
    if (ctx->pstream_persistent) {
        unsigned* c = nullptr;
        HRT_CUDA(cudaHostAlloc(&c, 32 * sizeof(unsigned), cudaHostAllocMapped));
        memset(c, 0, 32 * sizeof(unsigned));
        ctx->pctrl = c;
        HRT_CUDA(cudaHostGetDevicePointer(&ctx->pctrl_d, c, 0));
        HRT_CUDA(cudaFuncSetAttribute(pos_stream_server_kernel<POS_FULL_BODY_POS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    ctx->pstream_open = true;
    return 0;
}

int hrt_stream_pos_frame(hrt_ctx* ctx, const float* h_body_t, const float* h_lhand_t, const float* h_rhand_t,
                         float* h_robot_local_q, float* h_dof) {
    return hrt_stream_pos_frame_ex(ctx, h_body_t, h_lhand_t, h_rhand_t, nullptr, h_robot_local_q, h_dof, nullptr);
}

int hrt_stream_pos_frame_ex(hrt_ctx* ctx, const float* h_body_t, const float* h_lhand_t, const float* h_rhand_t,
                            const float* h_body_q, float* h_robot_local_q, float* h_dof, float* h_body_gq) {
    if (!ctx || !ctx->pstream_open) return fail(HRT_E_NOT_CONFIGURED, "hrt_stream_pos_open has not been called");
    DeviceGuard dev_guard_;                      // the launch on ctx->pss needs the context's device current
    if (!ctx->pstream_persistent) {
        int rc = dev_guard_.enter(ctx);
        if (rc) return rc;
    }
    const PosArgs& a = ctx->pstream_args;
    if (!h_body_t || (a.lhand_t && (!h_lhand_t || !h_rhand_t)) || (a.body_q && !h_body_q)) return fail(HRT_E_INVALID_ARG, "null input");
    if (h_body_gq && !a.out_body_gq) return fail(HRT_E_INVALID_ARG, "stream opened without HRT_STREAM_BODY_GQ");
    const PosParams& pp = ctx->pos[ctx->pstream_mode];
    const size_t body_w = (size_t)(pp.n_body * 3 + 3) / 4 * 4, hand_w = (size_t)(pp.n_hand * 3 + 3) / 4 * 4;
    memcpy(ctx->pmb_in, h_body_t, (size_t)pp.n_body * 12);
    if (a.lhand_t) {
        memcpy(ctx->pmb_in + body_w, h_lhand_t, (size_t)pp.n_hand * 12);
        memcpy(ctx->pmb_in + body_w + hand_w, h_rhand_t, (size_t)pp.n_hand * 12);
    }
    if (a.body_q) memcpy(ctx->pmb_in + body_w + 2 * hand_w, h_body_q, (size_t)pp.n_bodyq * 16);
    if (!ctx->pstream_persistent) {
        int rc = launch_pos(ctx, ctx->pstream_mode, ctx->pstream_args, ctx->pss, 1);
        if (rc) return rc;
        HRT_CUDA(cudaStreamSynchronize(ctx->pss));
    } else {
        int rc = server_roundtrip(ctx->pctrl, &ctx->pseq, &ctx->pserver_launched, ctx->pss,
                                  [&](unsigned served) { return launch_pos_server(ctx, served); });
        if (rc) return rc;
    }
    if (h_robot_local_q) memcpy(h_robot_local_q, ctx->pmb_out, (size_t)pp.J_rob * 16);
    if (h_dof) memcpy(h_dof, ctx->pmb_out + pp.J_rob * 4, (size_t)(pp.J_rob - 1) * 4);
    if (h_body_gq) memcpy(h_body_gq, ctx->pmb_out + pp.J_rob * 4 + 32, (size_t)pp.J_bq * 16);
    return 0;
}

int hrt_stream_pos_close(hrt_ctx* ctx) {
    if (!ctx || !ctx->pstream_open) return 0;
    DeviceGuard dev_guard_;
    dev_guard_.enter(ctx);
    if (ctx->pctrl) ctx->pctrl[1] = 1u;               // tell a resident server to leave
    if (ctx->pss) { cudaStreamSynchronize(ctx->pss); cudaStreamDestroy(ctx->pss); ctx->pss = nullptr; }
    if (ctx->pmb_in) { cudaFreeHost(ctx->pmb_in); ctx->pmb_in = nullptr; }
    if (ctx->pmb_out) { cudaFreeHost(ctx->pmb_out); ctx->pmb_out = nullptr; }
    if (ctx->pctrl) { cudaFreeHost(const_cast<unsigned*>(ctx->pctrl)); ctx->pctrl = nullptr; ctx->pctrl_d = nullptr; }
    ctx->pstream_open = false;
    ctx->pserver_launched = false;
    return 0;
}

int hrt_rescale_motion(hrt_ctx* ctx, int tree, int64_t B, const float* d_gt, const float* dir3, float* d_out, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (B == 0) return 0;
    if (B < 0 || !d_gt || !d_out) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    RescaleParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.J = t->tp.J;
    for (int j = 0; j < rp.J; ++j) {
        rp.parent[j] = t->tp.parent[j];
        for (int k = 0; k < 3; ++k) rp.off[j][k] = t->tp.jr[j].off[k];
    }
    for (int k = 0; k < 3; ++k) rp.dir[k] = dir3 ? dir3[k] : 1.f;
    const size_t smem = (size_t)MOT_WARPS * 2 * 32 * rp.J * 3 * sizeof(float);
    const long long tiles = (B + 31) / 32;
    static LaunchCache rs_cache;                            // grid = the CTAs resident together (it was a fixed 2 per SM)
    int grid = 1;
    if ((rc = grid_for(ctx, rescale_motion_kernel, MOT_WARPS * 32, smem, (tiles + MOT_WARPS - 1) / MOT_WARPS, &grid, &rs_cache))) return rc;
    rescale_motion_kernel<<<grid, MOT_WARPS * 32, smem, (cudaStream_t)stream>>>(rp, d_gt, B, d_out);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_rebuild_global_rotation(hrt_ctx* ctx, int tree, int64_t B, const float* d_gt, int n_kabsch,
                                const int32_t* kabsch_joint, const int32_t* kabsch_pts, float* d_out_gq, void* stream) {
    HRT_ENTER(ctx);
    Tree* t;
    if ((rc = get_tree(ctx, tree, &t))) return rc;
    if (n_kabsch < 0 || n_kabsch > 2 || (n_kabsch && (!kabsch_joint || !kabsch_pts)))
        return fail(HRT_E_INVALID_ARG, "n_kabsch must be 0..2 with its tables");
    if (B == 0) return 0;
    if (B < 0 || !d_gt || !d_out_gq) return fail(HRT_E_INVALID_ARG, "bad B / null pointer");
    if (!aligned16(d_out_gq)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    RebuildParams rp;
    memset(&rp, 0, sizeof(rp));
    const int J = rp.J = t->tp.J;
    for (int j = 0; j < J; ++j) {
        rp.parent[j] = t->tp.parent[j];
        for (int k = 0; k < 3; ++k) rp.off[j][k] = t->tp.jr[j].off[k];
    }
    rp.n_kabsch = n_kabsch;
    for (int k = 0; k < n_kabsch; ++k) {
        if (kabsch_joint[k] < 0 || kabsch_joint[k] >= J) return fail(HRT_E_INVALID_ARG, "kabsch_joint[%d] out of range", k);
        rp.kabsch_joint[k] = kabsch_joint[k];
        for (int n = 0; n < 3; ++n) {
            const int pj = kabsch_pts[k * 3 + n];
            if (pj < 0 || pj >= J) return fail(HRT_E_INVALID_ARG, "kabsch_pts[%d][%d] out of range", k, n);
            rp.kabsch_pts[k][n] = pj;
        }
    }
    // main.py:146: skip the root and every child of a Kabsch-fitted joint
    for (int j = 0; j < J; ++j) {
        bool skip = j == 0 || rp.parent[j] < 0;
        for (int k = 0; k < n_kabsch; ++k) skip |= rp.parent[j] == rp.kabsch_joint[k];
        rp.skip[j] = skip ? 1 : 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    HRT_CUDA(cudaMemsetAsync(ctx->d_scalars, 0, HRT_MAX_JOINTS * sizeof(unsigned), st));
    {
        const int grid = (int)std::max(1LL, std::min(((long long)B + 255) / 256, (long long)ctx->sm_count * 8));
        bone_max_norm_kernel<<<grid, 256, 0, st>>>(rp, d_gt, B, ctx->d_scalars);
        HRT_CUDA(cudaGetLastError());
    }
    const size_t smem = (size_t)MOT_WARPS * 32 * J * 7 * sizeof(float);
    const long long tiles = (B + 31) / 32;
    // grid = the CTAs resident together (shared memory allows 3 per SM for 21 joints; it was a fixed 2)
    static LaunchCache rc_cache;
    int grid = 1;
    if ((rc = grid_for(ctx, rebuild_rotation_kernel, MOT_WARPS * 32, smem, (tiles + MOT_WARPS - 1) / MOT_WARPS, &grid, &rc_cache))) return rc;
    rebuild_rotation_kernel<<<grid, MOT_WARPS * 32, smem, st>>>(rp, d_gt, B, ctx->d_scalars, d_out_gq);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

// numpy's pairwise summation (what phi_x.sum() runs): n < 8 sequential; n <= 128 eight partial sums; else split
static double np_pairwise_sum(const double* a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// scipy.ndimage._filters._gaussian_kernel1d(sigma, order=0, radius=int(4*sigma+0.5)): exp(-0.5/sigma^2 * x^2) / sum
static int gauss_params(GaussParams* gp, double sigma) {
    const int radius = (int)(4.0 * sigma + 0.5);
    if (!(sigma > 0.0) || radius > HRT_GAUSS_MAX_RADIUS) return -1;
    gp->radius = radius;
    const int n = 2 * radius + 1;
    for (int k = 0; k < n; ++k) {
        const double x = (double)(k - radius);
        gp->w[k] = std::exp(-0.5 / (sigma * sigma) * (x * x));
    }
    const double sum = np_pairwise_sum(gp->w, n);
    for (int k = 0; k < n; ++k) gp->w[k] /= sum;
    return 0;
}

static int ew_grid(hrt_ctx* ctx, long long n) { return (int)std::max(1LL, std::min((n + 255) / 256, (long long)ctx->sm_count * 16)); }

// sigma = 2 (radius 8): the sliding-window form, one thread per (run of GAUSS_RUN frames, component)
static void launch_gauss_sigma2(const GaussParams& gp, const float* x, long long T, long long C, float* out, cudaStream_t st) {
    const long long threads = (T + GAUSS_RUN - 1) / GAUSS_RUN * C;
    gauss_filter_frames_window_kernel<float, float, 8><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(gp, x, T, C, out);
}

int hrt_motion_velocity(hrt_ctx* ctx, int64_t T, int64_t J, const float* d_gt, float dt, int gaussian,
                        float* d_scratch, float* d_out, void* stream) {
    HRT_ENTER(ctx);
    if (T < 2) return fail(HRT_E_INVALID_ARG, "np.gradient needs at least 2 frames");
    if (J < 1 || !d_gt || !d_out || (gaussian && !d_scratch)) return fail(HRT_E_INVALID_ARG, "bad J / null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long C = J * 3, n = T * C;
    frame_gradient_kernel<<<ew_grid(ctx, n), 256, 0, st>>>(d_gt, T, C, dt, gaussian ? d_scratch : d_out);
    HRT_CUDA(cudaGetLastError());
    if (gaussian) {
        GaussParams gp;
        gauss_params(&gp, 2.0);
        launch_gauss_sigma2(gp, d_scratch, T, C, d_out, st);
        HRT_CUDA(cudaGetLastError());
    }
    return 0;
}

int hrt_motion_angular_velocity(hrt_ctx* ctx, int64_t T, int64_t J, const float* d_gq, float dt, int gaussian,
                                float* d_scratch, float* d_out, void* stream) {
    HRT_ENTER(ctx);
    if (T < 1 || J < 1 || !d_gq || !d_out || (gaussian && !d_scratch)) return fail(HRT_E_INVALID_ARG, "bad T / J / null pointer");
    if (!aligned16(d_gq)) return fail(HRT_E_ALIGNMENT, "buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = T * J;
    angular_velocity_raw_kernel<<<ew_grid(ctx, n), 256, 0, st>>>(reinterpret_cast<const float4*>(d_gq), T, J, dt, gaussian ? d_scratch : d_out);
    HRT_CUDA(cudaGetLastError());
    if (gaussian) {
        GaussParams gp;
        gauss_params(&gp, 2.0);
        launch_gauss_sigma2(gp, d_scratch, T, J * 3, d_out, st);
        HRT_CUDA(cudaGetLastError());
    }
    return 0;
}

int hrt_forward_vector(hrt_ctx* ctx, int64_t T, int64_t J, const float* d_gt, int left_shoulder, int right_shoulder,
                       int left_hip, int right_hip, double sigma, double* d_scratch, double* d_out, void* stream) {
    HRT_ENTER(ctx);
    if (T < 0 || J < 1) return fail(HRT_E_INVALID_ARG, "bad T / J");
    if (T == 0) return 0;
    if (!d_gt || !d_scratch || !d_out) return fail(HRT_E_INVALID_ARG, "null pointer");
    const int idx[4] = {left_shoulder, right_shoulder, left_hip, right_hip};
    for (int k = 0; k < 4; ++k)
        if (idx[k] < 0 || idx[k] >= J) return fail(HRT_E_INVALID_ARG, "joint index out of range");
    GaussParams gp;
    if (gauss_params(&gp, sigma)) return fail(HRT_E_INVALID_ARG, "gaussian width must be in (0, 32)");
    cudaStream_t st = (cudaStream_t)stream;
    forward_raw_kernel<<<ew_grid(ctx, T), 256, 0, st>>>(d_gt, T, (int)J, idx[0], idx[1], idx[2], idx[3], d_scratch);
    HRT_CUDA(cudaGetLastError());
    gauss_filter_frames_kernel<double, double><<<ew_grid(ctx, T * 3), 256, 0, st>>>(gp, d_scratch, T, 3, d_out);
    HRT_CUDA(cudaGetLastError());
    normalize_rows3_f64_kernel<<<ew_grid(ctx, T), 256, 0, st>>>(d_out, T);
    HRT_CUDA(cudaGetLastError());
    return 0;
}

extern "C++" {
namespace {
template <int OP>
int launch_rot_op(hrt_ctx* ctx, const RotOpArgs& a, cudaStream_t st) {
    const long long tiles = (a.n + 31) / 32;
    // one wave of exactly the CTAs that are resident together (grid-stride loop inside): a fixed 8 per SM left a second,
    // one-third-full wave behind for every instantiation above 32 registers
    static int occ[16] = {0};
    const int d = ctx->device & 15;
    if (!occ[d]) {
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, rot_op_kernel<OP>, ROT_WARPS * 32, 0) != cudaSuccess || o < 1) o = 4;
        occ[d] = o;
    }
    const int grid = (int)std::max(1LL, std::min((tiles + ROT_WARPS - 1) / ROT_WARPS, (long long)ctx->sm_count * occ[d]));
    rot_op_kernel<OP><<<grid, ROT_WARPS * 32, 0, st>>>(a);
    return 0;
}
template <int OP>
void rot_op_info(int* ni, int* wi, int* no, int* wo) {
    using Tr = RotOpTraits<OP>;
    *ni = Tr::NI; *no = Tr::NO;
    for (int k = 0; k < 4; ++k) wi[k] = Tr::WI[k];
    for (int k = 0; k < 3; ++k) wo[k] = Tr::WO[k];
}
#define HRT_FOR_EACH_OP(X)                                                                                          \
    X(OP_QUAT_MUL) X(OP_QUAT_MUL_NORM) X(OP_QUAT_MUL_THREE) X(OP_QUAT_MUL_FOUR) X(OP_QUAT_POS) X(OP_QUAT_ABS)       \
    X(OP_QUAT_UNIT) X(OP_QUAT_NORMALIZE) X(OP_QUAT_CONJUGATE) X(OP_QUAT_ROTATE) X(OP_QUAT_FROM_ANGLE_AXIS)          \
    X(OP_QUAT_FROM_ROTATION_MATRIX) X(OP_QUAT_ANGLE_AXIS) X(OP_QUAT_YAW_ROTATION) X(OP_TRANSFORM_INVERSE)           \
    X(OP_TRANSFORM_MUL) X(OP_TRANSFORM_APPLY) X(OP_ROT_MATRIX_DET) X(OP_ROT_MATRIX_FROM_QUATERNION)                 \
    X(OP_PROJECT_QUAT_TO_AXIS) X(OP_EXTRACT_ROTATION_ALONG_AXIS) X(OP_NORMALIZE_ANGLE) X(OP_QUAT_TO_ANGLE_AXIS)     \
    X(OP_QUAT_TO_EXP_MAP) X(OP_EXP_MAP_TO_ANGLE_AXIS) X(OP_EXP_MAP_TO_QUAT) X(OP_ANGLE_AXIS_TO_EXP_MAP)             \
    X(OP_QUAT_BETWEEN_TWO_VECS) X(OP_PROJ_IN_PLANE) X(OP_RADIANS_BETWEEN_VECS) X(OP_QUAT_SLERP)                     \
    X(OP_QUAT_TO_DOF_POS) X(OP_EULER_SPLIT) X(OP_EULER_ANGLES_F64) X(OP_COORD_TRANSFORM)                            \
    X(OP_CAL_SHOULDER_PR) X(OP_CAL_ELBOWP_SHOULDERY)
}  // namespace
}  // extern "C++"

int hrt_rot_op_info(int op, int* n_in, int* in_width4, int* n_out, int* out_width3) {
    if (!n_in || !in_width4 || !n_out || !out_width3) return fail(HRT_E_INVALID_ARG, "null pointer");
    switch (op) {
#define X(OP) case OP: rot_op_info<OP>(n_in, in_width4, n_out, out_width3); return 0;
        HRT_FOR_EACH_OP(X)
#undef X
    }
    return fail(HRT_E_INVALID_ARG, "unknown op %d", op);
}

int hrt_rot_op(hrt_ctx* ctx, int op, int64_t n, const float* const* d_in4, const int64_t* period4, int iparam,
               float fparam, float* const* d_out3, void* stream) {
    HRT_ENTER(ctx);
    int ni, wi[4], no, wo[3];
    if ((rc = hrt_rot_op_info(op, &ni, wi, &no, wo))) return rc;
    if (n < 0 || !d_in4 || !period4 || !d_out3) return fail(HRT_E_INVALID_ARG, "bad n / null pointer");
    if (n == 0) return 0;
    RotOpArgs a;
    memset(&a, 0, sizeof(a));
    a.n = n; a.iparam = iparam; a.fparam = fparam;
    for (int k = 0; k < ni; ++k) {
        if (!d_in4[k]) return fail(HRT_E_INVALID_ARG, "operand %d is null", k);
        if (period4[k] < 0) return fail(HRT_E_INVALID_ARG, "period[%d] < 0", k);
        a.in[k] = d_in4[k];
        a.period[k] = period4[k];
    }
    for (int k = 0; k < no; ++k) {
        if (!d_out3[k]) return fail(HRT_E_INVALID_ARG, "result %d is null", k);
        a.out[k] = d_out3[k];
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (op) {
#define X(OP) case OP: launch_rot_op<OP>(ctx, a, st); break;
        HRT_FOR_EACH_OP(X)
#undef X
    }
    HRT_CUDA(cudaGetLastError());
    return 0;
}

int hrt_max_norm3(hrt_ctx* ctx, int64_t n, const float* d_v, float* h_out, void* stream) {
    HRT_ENTER(ctx);
    if (n < 0 || !h_out || (n > 0 && !d_v)) return fail(HRT_E_INVALID_ARG, "bad n / null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned* slot = ctx->d_scalars + HRT_MAX_JOINTS;
    HRT_CUDA(cudaMemsetAsync(slot, 0, sizeof(unsigned), st));
    if (n > 0) {
        max_norm3_kernel<<<ew_grid(ctx, n), 256, 0, st>>>(d_v, n, slot);
        HRT_CUDA(cudaGetLastError());
    }
    unsigned bits = 0;
    HRT_CUDA(cudaMemcpyAsync(&bits, slot, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    HRT_CUDA(cudaStreamSynchronize(st));
    memcpy(h_out, &bits, 4);
    return 0;
}

int hrt_cal_joint_quat(hrt_ctx* ctx, int64_t n, int n_points, const float* d_zero, int64_t zero_period,
                       const float* d_motion, float* d_out_q, void* stream) {
    HRT_ENTER(ctx);
    if (n < 0 || n_points < 1 || zero_period < 0) return fail(HRT_E_INVALID_ARG, "bad n / n_points / zero_period");
    if (n == 0) return 0;
    if (!d_zero || !d_motion || !d_out_q) return fail(HRT_E_INVALID_ARG, "null pointer");
    if (!aligned16(d_out_q)) return fail(HRT_E_ALIGNMENT, "d_out_q must be 16-byte aligned");
    const int grid = (int)std::max(1LL, std::min(((long long)n + 127) / 128, (long long)ctx->sm_count * 8));
    kabsch_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_zero, zero_period, d_motion, n, n_points, d_out_q);
    HRT_CUDA(cudaGetLastError());
    return 0;
}


}  // extern "C"
