// sqk_api.cu -- the C ABI of libsqk.so (include/sqk.h): context, scratch memory, the
// host-buffer streaming pipeline and the device-buffer enqueue path around the three kernels
// (sqk_stats.cuh, sqk_dtw.cuh, sqk_segmenter.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "sqk_dtw_launch.cuh"
#include "sqk_dtw_lb.cuh"
#include "sqk_segmenter.cuh"
#include "sqk_adapter.cuh"
#include "sqk_f64.cuh"
#include "sqk_stats.cuh"
#include "sqk_stats2.cuh"
#include "sqk_stats3.cuh"
#include "sqk_rollmean.cuh"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? SQK_ERR_NOMEM : SQK_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                   \
    } while (0)

#define TRY(call)                  \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != SQK_OK) return rc_; \
    } while (0)

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

static int ensure(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return SQK_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        want = bytes;
        e = cudaMalloc(&b.p, want);
    }
    if (e != cudaSuccess) { b.p = nullptr; return fail(SQK_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(e)); }
    b.cap = want;
    return SQK_OK;
}

static void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
}

struct HostBuf {              // pinned host staging
    void *p = nullptr;
    size_t cap = 0;
};

static int ensure_host(HostBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return SQK_OK;
    if (b.p) { cudaFreeHost(b.p); b.p = nullptr; b.cap = 0; }
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaHostAlloc(&b.p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { b.p = nullptr; cudaGetLastError(); return fail(SQK_ERR_NOMEM, "cudaHostAlloc(%zu bytes): %s", want, cudaGetErrorString(e)); }
    b.cap = want;
    return SQK_OK;
}

struct Pending { void *dst; const void *src; size_t bytes; };

struct Slot {                 // everything one in-flight chunk needs
    cudaStream_t stream = nullptr;
    DevBuf signals, offsets, stats, hits, nkept, segs, nsegs, counter, gstage, pa_off, pa_scale, ynorm, codes;
    DevBuf jobs, fbjobs, lbreads, jobres;   // two-pass DTW plan (sqk_dtw_plan.cuh)
    DevBuf redo, mask, rm_p, rm_masks, bnd_a, bnd_b, jobs2, rtjobs, pending;        // sqk_stats2_kernel: redo list ([0] = length, entries from [4]); segmenter bit masks
    HostBuf hin;              // a PAGEABLE caller buffer is brought here (pinned) by all host threads before its H2D copy
    HostBuf hout[2];          // results land here (pinned) so the D2H copy never blocks the host ...
    Pending pend[2];          // ... and move to the caller's (possibly pageable) arrays when the slot is recycled
    int n_pend = 0;
};

static bool is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

extern "C" void sqk_parallel_memcpy(void *dst, const void *src, size_t bytes, int n_threads);   // sqk_tsv.cpp (OpenMP)

// caller's host samples -> device.  A pinned source goes out as one asynchronous copy.  A pageable one would be staged by
// the driver through its own pinned buffer on ONE thread (~10 GB/s: the call is then 4-5x slower than the PCIe bound); it
// is copied into the slot's pinned staging buffer by all host threads instead (the slot's previous chunk is done: recycle),
// and that copy overlaps the other slot's transfers and kernels.
static int signals_to_device(Slot &s, cudaStream_t st, void *dst, const void *src, size_t bytes, bool src_pinned)
{
    if (bytes == 0) return SQK_OK;
    if (src_pinned) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return SQK_OK;
    }
    TRY(ensure_host(s.hin, bytes));
    sqk_parallel_memcpy(s.hin.p, src, bytes, 0);
    CU(cudaMemcpyAsync(dst, s.hin.p, bytes, cudaMemcpyHostToDevice, st));
    return SQK_OK;
}

// device -> caller's host array, without stalling the pipeline when that array is pageable
static int result_to_host(Slot &s, int which, void *dst, const void *dev_src, size_t bytes, bool dst_pinned)
{
    if (bytes == 0) return SQK_OK;
    if (dst_pinned) {
        CU(cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToHost, s.stream));
        return SQK_OK;
    }
    TRY(ensure_host(s.hout[which], bytes));
    CU(cudaMemcpyAsync(s.hout[which].p, dev_src, bytes, cudaMemcpyDeviceToHost, s.stream));
    s.pend[s.n_pend++] = Pending{dst, s.hout[which].p, bytes};
    return SQK_OK;
}

static int recycle(Slot &s)   // wait for the slot's previous chunk and hand its results over
{
    CU(cudaStreamSynchronize(s.stream));
    for (int i = 0; i < s.n_pend; i++) memcpy(s.pend[i].dst, s.pend[i].src, s.pend[i].bytes);
    s.n_pend = 0;
    return SQK_OK;
}

struct TimeRec { int kid; cudaEvent_t a, b; };

struct sqk_ctx {
    int device = 0;
    int n_sms = 0;
    int smem_optin = 0;
    int clock_khz = 0;
    int l2_bytes = 0;
    int cc = 0;
    cudaStream_t user_stream = nullptr;
    bool use_user_stream = false;
    Slot slot[2];
    DevBuf model;
    DevBuf scratch8;
    bool timing = false;
    std::vector<TimeRec> recs;
    std::vector<cudaEvent_t> pool;
    sqk_timing acc{};
    int force_lanes = 0;
    int dtw_plan = SQK_PLAN_AUTO;
    int lb_want_k = 12;   // measured on B200: K=10/L=8 5.85 ms vs K=20/L=4 6.11 ms per 100k x 4096 x 80
    int64_t chunk_samples = 0;     // host mode: samples per in-flight chunk (0 = default / SQK_CHUNK_SAMPLES)
    int stats_smem_set32 = -1, stats_smem_set128 = -1, stats_smem_set256 = -1;
    int stats2_smem_set[5] = {-1, -1, -1, -1, -1};
    int stats3_smem_set[4] = {-1, -1, -1, -1};
    int stats_gen = 0;                // 0 = automatic, 1 = first-generation kernel only (experiments / tests)
    // device-mode calls share the ctx scratch in stream order: the end of every enqueue is marked with an event and a
    // call that arrives on a different stream waits for it first
    PeerOut peers{};                  // multi-GPU publication targets of device-mode sqk_motifseq (sqk_ctx_set_hit_peers)
    int64_t peer_base = 0;            // first record of this rank in a gathered buffer
    int64_t peer_cap = 0;             // records a gathered buffer holds (0 = not told: no bounds check)
    PeerFlags flags{};
    int64_t n_launches = 0;           // kernels launched (sqk_ctx_get_launches)
    cudaEvent_t dev_done = nullptr;
    cudaStream_t dev_last = nullptr;
    bool dev_any = false;
    std::vector<double> model_host;   // what c->model holds on the device (uploads are skipped when nothing changed)
    HostBuf model_stage;              // pinned staging of the model upload (no implicit sync on the caller's stream)
};

struct Guard {   // make the ctx device current for the duration of a call
    int prev = -1;
    bool ok = false;
    explicit Guard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Host-mode pipelines: when a call fails part-way, earlier chunks may still be copying into the caller's result arrays
// and their pinned->pageable hand-overs are still queued.  The caller is about to see an error (and may free those
// arrays): wait for what is in flight and DROP the queued hand-overs instead of leaving them for the next call.
struct PipeGuard {
    sqk_ctx *c;
    bool armed = true;
    explicit PipeGuard(sqk_ctx *ctx) : c(ctx) {}
    ~PipeGuard()
    {
        if (!armed) return;
        for (int i = 0; i < 2; i++) {
            cudaStreamSynchronize(c->slot[i].stream);
            c->slot[i].n_pend = 0;
        }
        cudaGetLastError();
    }
};

// Device-mode calls use ctx-level scratch (slot 0, the model buffer): order a call on stream `st` behind the previous
// device-mode call if that one ran on another stream, and behind host-mode work on the slot streams.
static int dev_begin(sqk_ctx *c, cudaStream_t st)
{
    if (c->dev_any && c->dev_last != st) CU(cudaStreamWaitEvent(st, c->dev_done, 0));
    return SQK_OK;
}

static int host_begin(sqk_ctx *c)   // host-mode pipelines run on the slot streams: wait for enqueued device-mode work
{
    if (c->dev_any) CU(cudaEventSynchronize(c->dev_done));
    return SQK_OK;
}

static int dev_end(sqk_ctx *c, cudaStream_t st)
{
    if (!c->dev_done) CU(cudaEventCreateWithFlags(&c->dev_done, cudaEventDisableTiming));
    CU(cudaEventRecord(c->dev_done, st));
    c->dev_last = st; c->dev_any = true;
    return SQK_OK;
}

// Models are a few hundred bytes: staged through pinned memory, and only when they differ from what the device holds.
static int upload_models(sqk_ctx *c, const double *models, size_t n_points, cudaStream_t st)
{
    const size_t bytes = n_points * sizeof(double);
    if (c->model.p && c->model_host.size() == n_points && memcmp(c->model_host.data(), models, bytes) == 0) return SQK_OK;
    // the previous upload (and every kernel reading the old model) must be done before the staging buffer is rewritten
    if (c->dev_any) CU(cudaEventSynchronize(c->dev_done));
    CU(cudaStreamSynchronize(c->slot[0].stream));
    CU(cudaStreamSynchronize(c->slot[1].stream));
    TRY(ensure(c->model, bytes));
    TRY(ensure_host(c->model_stage, bytes));
    memcpy(c->model_stage.p, models, bytes);
    c->model_host.clear();
    CU(cudaMemcpyAsync(c->model.p, c->model_stage.p, bytes, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    c->model_host.assign(models, models + n_points);
    return SQK_OK;
}

static int tick(sqk_ctx *c, int kid, cudaStream_t st, cudaEvent_t *b_out)
{
    *b_out = nullptr;
    if (!c->timing) return SQK_OK;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; i++) {
        if (!c->pool.empty()) { ev[i] = c->pool.back(); c->pool.pop_back(); }
        else CU(cudaEventCreate(&ev[i]));
    }
    CU(cudaEventRecord(ev[0], st));
    c->recs.push_back({kid, ev[0], ev[1]});
    *b_out = ev[1];
    return SQK_OK;
}

static int tock(cudaEvent_t b, cudaStream_t st)
{
    if (b) CU(cudaEventRecord(b, st));
    return SQK_OK;
}

static int drain_timing(sqk_ctx *c)
{
    for (auto &r : c->recs) {
        CU(cudaEventSynchronize(r.b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, r.a, r.b));
        c->acc.launches[r.kid] += 1;
        c->acc.ms[r.kid] += ms;
        c->pool.push_back(r.a);
        c->pool.push_back(r.b);
    }
    c->recs.clear();
    return SQK_OK;
}

// ------------------------------------------------------------------------------------------
// small helper kernels
// ------------------------------------------------------------------------------------------
__global__ void sqk_max_len_kernel(const int64_t *offsets, int64_t n_reads, unsigned long long *out)
{
    unsigned long long m = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_reads; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t d = offsets[i + 1] - offsets[i];
        if (d > 0 && (unsigned long long)d > m) m = (unsigned long long)d;
        if (d < 0) m = ~0ull;   // offsets not monotone
    }
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(SQK_FULL_MASK, m, d);
        m = o > m ? o : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// Debug / plotting path (MotifSeq.py:447 `-x`, :507-509 cost[-1,]): one read, one CTA.  Writes the
// normalised post-outlier signal and the whole last DTW row.  Row i is owned by thread i; columns
// advance as an anti-diagonal through shared memory -- an implementation independent of
// sqk_dtw_kernel, which the tests also use to cross-check it.
__global__ void sqk_trace_normalise_kernel(const int16_t *sig, int64_t n, int lo, int hi, double center, double scale,
                                           double *out, int *n_out)
{
    __shared__ int warp_tot[32];
    __shared__ int running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (tid == 0) running = 0;
    __syncthreads();
    for (int64_t b = 0; b < n; b += blockDim.x) {
        const int64_t i = b + tid;
        const int v = i < n ? (int)sig[i] : 0;
        const bool keep = i < n && v > lo && v < hi;
        const unsigned bal = __ballot_sync(SQK_FULL_MASK, keep);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int pos = running + __popc(bal & ((1u << lane) - 1u));
        int all = 0;
        for (int w = 0; w < nw; w++) { if (w < warp) pos += warp_tot[w]; all += warp_tot[w]; }
        if (keep) out[pos] = __ddiv_rn(__dsub_rn((double)v, center), scale);
        __syncthreads();
        if (tid == 0) running += all;
        __syncthreads();
    }
    if (tid == 0) *n_out = running;
}

__global__ void sqk_trace_lastrow_kernel(const double *y, int m, const double *x, int n, double *last_row)
{
    extern __shared__ double sh[];   // cur[n], prev1[n], prev2[n]: columns j, j-1 (per row, skewed)
    double *col1 = sh, *col2 = sh + n;
    const int i = threadIdx.x;
    const double xi = i < n ? x[i] : 0.0;
    double left = SQK_INF_D;         // C[i][j-1]
    if (i < n) { col1[i] = SQK_INF_D; col2[i] = SQK_INF_D; }
    __syncthreads();
    for (int t = 0; t < m + n - 1; t++) {
        const int j = t - i;
        double nc = SQK_INF_D;
        const bool on = (i < n && j >= 0 && j < m);
        if (on) {
            // col1[i-1] = C[i-1][j] (written by thread i-1 last step), col2[i-1] = C[i-1][j-1]
            double best;
            if (i == 0) best = 0.0;
            else {
                const double up = col1[i - 1], dg = col2[i - 1];
                best = up;
                if (dg < best) best = dg;
                if (left < best) best = left;
                if (j == 0) best = up;
            }
            nc = __dadd_rn(fabs(__dsub_rn(xi, y[j])), best);
        }
        __syncthreads();
        if (on) {
            col2[i] = col1[i];
            col1[i] = nc;
            left = nc;
            if (i == n - 1) last_row[j] = nc;
        } else if (i < n && j >= m) {
            col2[i] = col1[i];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// launch plumbing shared by host and device mode
// ------------------------------------------------------------------------------------------
struct View {                 // a set of reads resident on the device
    const int16_t *base;      // base[i] = absolute sample i
    int64_t alloc_lo, alloc_hi;
    const int64_t *offsets;   // indexable by absolute read id
    int64_t read0;
    int64_t n_reads;
    int64_t max_len;
};

static inline int clamp_lim(int v) { return v < -40000 ? -40000 : (v > 40000 ? 40000 : v); }   // samples are int16

template <int NT>
static int launch_stats_nt(sqk_ctx *c, Slot &s, cudaStream_t st, StatsArgs &a, const View &v, int *smem_set,
                           bool timed = true, int64_t max_grid = 0)
{
    constexpr int TPB = StatsCta<NT>::threads;
    constexpr int GROUPS = TPB / NT;
    const size_t fixed = (sizeof(StatsShared) + 15) & ~(size_t)15;
    // staging capacity per group: the longest read if it fits (2 bytes/sample)
    const int64_t budget = ((int64_t)c->smem_optin - 1024) / GROUPS - (int64_t)fixed;
    int64_t cap_max = budget / 2;
    cap_max &= ~127ll;
    int64_t cap = (std::max<int64_t>(v.max_len, 8) + 127) & ~127ll;
    if (cap > cap_max) cap = cap_max;
    const size_t group_bytes = fixed + (size_t)cap * 2;
    const int dyn = (int)(group_bytes * GROUPS);
    if (dyn > *smem_set) {
        CU(cudaFuncSetAttribute(sqk_stats_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        *smem_set = dyn;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sqk_stats_kernel<NT>, TPB, dyn));
    if (per_sm < 1) per_sm = 1;
    const int64_t want = (v.n_reads + GROUPS - 1) / GROUPS;
    int64_t grid = std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->n_sms * per_sm));
    if (max_grid > 0) grid = std::min(grid, max_grid);
    a.cap = (int)cap; a.gstage = nullptr; a.gstage_stride = 0;
    if (v.max_len > cap && a.mode != SQK_STATS_NONE) {
        const int64_t stride = (v.max_len + 7) & ~7ll;
        TRY(ensure(s.gstage, (size_t)grid * GROUPS * stride * sizeof(int16_t)));
        a.gstage = (int16_t *)s.gstage.p; a.gstage_stride = stride;
    }
    cudaEvent_t eb = nullptr;
    if (timed) TRY(tick(c, SQK_K_STATS, st, &eb));
    sqk_stats_kernel<NT><<<(unsigned)grid, TPB, dyn, st>>>(a);
    CU(cudaGetLastError());
    c->n_launches++;
    if (timed) TRY(tock(eb, st));
    return SQK_OK;
}

// What the statistics pass left behind for the kernels after it.
struct StatsOut {
    bool v2 = false;                  // sqk_stats2_kernel ran; reads on the redo list were done by sqk_stats_kernel
    const int *redo = nullptr;        // redo list (launch-local read indices) ...
    const unsigned *n_redo = nullptr; // ... and its length, both in device memory
    const uint32_t *mask = nullptr;   // segmenter mode: in-range bit masks [n_reads][mask_stride] (not for redo reads)
    int mask_stride = 0;
};

// K1, second generation (sqk_stats2.cuh) for reads that fit its shared-memory window; the reads it hands back are
// worked off by the first-generation kernel through the redo list.
static int launch_stats2(sqk_ctx *c, Slot &s, cudaStream_t st, StatsArgs &a, const View &v, bool want_mask, StatsOut *so)
{
    Stats2Args A{};
    const int cap_units = (int)((std::max<int64_t>(v.max_len, 8) + 14 + 7) / 8);
    const int nseg = (cap_units + 15) / 16;
    A.buf_bytes = nseg * 256;
    A.hist_words = a.mode == SQK_STATS_MEDMAD ? 2 * SQK_S2_MAX_BINS : (a.mode == SQK_STATS_SEGMENTER ? SQK_S2_MAX_BINS : 0);
    A.mask_words = want_mask ? ((nseg * 4 + 2 + 3) & ~3) : 0;
    A.mask_stride = 0;
    if (want_mask) {
        A.mask_stride = (int)((((v.max_len + 31) / 32) + 3) & ~3ll);
        if (A.mask_stride < 4) A.mask_stride = 4;
        TRY(ensure(s.mask, (size_t)v.n_reads * A.mask_stride * sizeof(uint32_t)));
        A.mask = (uint32_t *)s.mask.p;
    }
    TRY(ensure(s.redo, ((size_t)v.n_reads + 4) * sizeof(int)));
    A.n_redo = (unsigned *)s.redo.p;
    A.redo = (int *)s.redo.p + 4;
    CU(cudaMemsetAsync(s.redo.p, 0, 16, st));
    const int dyn = (int)(((sizeof(S2Shared) + 15) & ~(size_t)15) + 4 * (size_t)(A.hist_words + A.mask_words) + 2 * (size_t)A.buf_bytes);
    void (*kern)(const Stats2Args) = nullptr;
    int which = 0;
    const bool pa = a.mode == SQK_STATS_SEGMENTER && a.pa_offset != nullptr;
    switch (a.mode) {
    case SQK_STATS_ZSCALE: kern = sqk_stats2_kernel<SQK_STATS_ZSCALE, false>; which = 0; break;
    case SQK_STATS_MEDMAD: kern = sqk_stats2_kernel<SQK_STATS_MEDMAD, false>; which = 1; break;
    case SQK_STATS_NONE: kern = sqk_stats2_kernel<SQK_STATS_NONE, false>; which = 2; break;
    default:
        if (pa) { kern = sqk_stats2_kernel<SQK_STATS_SEGMENTER, true>; which = 4; }
        else { kern = sqk_stats2_kernel<SQK_STATS_SEGMENTER, false>; which = 3; }
        break;
    }
    if (dyn > c->stats2_smem_set[which]) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        c->stats2_smem_set[which] = dyn;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SQK_S2_THREADS, dyn));
    if (per_sm < 1) per_sm = 1;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(v.n_reads, (int64_t)c->n_sms * per_sm));
    A.s = a;
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_STATS, st, &eb));
    kern<<<(unsigned)grid, SQK_S2_THREADS, dyn, st>>>(A);
    CU(cudaGetLastError());
    c->n_launches++;
    // the reads it handed back (more outliers than its exception list holds, windows wider than its histogram)
    a.list = A.redo; a.n_list = A.n_redo; a.extra_flags = SQK_FLAG_NO_MASK;
    so->v2 = true; so->redo = A.redo; so->n_redo = A.n_redo; so->mask = A.mask; so->mask_stride = A.mask_stride;
    const int rc = launch_stats_nt<128>(c, s, st, a, v, &c->stats_smem_set128, /*timed=*/false, /*max_grid=*/c->n_sms);
    TRY(tock(eb, st));
    return rc;
}

// K1, third generation (sqk_stats3.cuh): one warp per read, for the zscale / segmenter (raw integer) / none modes; the
// reads it hands back are worked off by the first-generation kernel through the redo list.
static int launch_stats3(sqk_ctx *c, Slot &s, cudaStream_t st, StatsArgs &a, const View &v, bool want_mask, StatsOut *so)
{
    Stats3Args A{};
    const int cap_units = (int)((std::max<int64_t>(v.max_len, 8) + 14 + 7) / 8);
    A.buf_bytes = ((cap_units + 15) / 16) * 256;
    const int lo1 = std::max(a.lo + 1, -32768), hi1 = std::min(a.hi - 1, 32767);
    const int nbins = hi1 >= lo1 ? hi1 - lo1 + 1 : 0;
    A.hist_words = (a.mode == SQK_STATS_SEGMENTER || a.mode == SQK_STATS_MEDMAD) ? std::max(128, (nbins + 127) & ~127) : 0;
    A.mask_words = want_mask ? ((A.buf_bytes / 64 + 2 + 3) & ~3) : 0;
    A.mask_stride = 0;
    if (want_mask) {
        A.mask_stride = (int)((((v.max_len + 31) / 32) + 3) & ~3ll);
        if (A.mask_stride < 4) A.mask_stride = 4;
        TRY(ensure(s.mask, (size_t)v.n_reads * A.mask_stride * sizeof(uint32_t)));
        A.mask = (uint32_t *)s.mask.p;
    }
    TRY(ensure(s.redo, ((size_t)v.n_reads + 4) * sizeof(int)));
    A.n_redo = (unsigned *)s.redo.p;
    A.redo = (int *)s.redo.p + 4;
    CU(cudaMemsetAsync(s.redo.p, 0, 16, st));
    const int dyn = (int)sqk_s3_smem_bytes(A);
    void (*kern)(const Stats3Args) = nullptr;
    int which = 0;
    switch (a.mode) {
    case SQK_STATS_ZSCALE: kern = sqk_stats3_kernel<SQK_STATS_ZSCALE>; which = 0; break;
    case SQK_STATS_NONE: kern = sqk_stats3_kernel<SQK_STATS_NONE>; which = 1; break;
    case SQK_STATS_MEDMAD: kern = sqk_stats3_kernel<SQK_STATS_MEDMAD>; which = 3; break;
    default: kern = sqk_stats3_kernel<SQK_STATS_SEGMENTER>; which = 2; break;
    }
    if (dyn > c->stats3_smem_set[which]) {
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
        c->stats3_smem_set[which] = dyn;
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, dyn));
    if (per_sm < 1) per_sm = 1;
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(v.n_reads, (int64_t)c->n_sms * per_sm));
    A.s = a;
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_STATS, st, &eb));
    kern<<<(unsigned)grid, 32, dyn, st>>>(A);
    CU(cudaGetLastError());
    c->n_launches++;
    a.list = A.redo; a.n_list = A.n_redo; a.extra_flags = SQK_FLAG_NO_MASK;
    so->v2 = true; so->redo = A.redo; so->n_redo = A.n_redo; so->mask = A.mask; so->mask_stride = A.mask_stride;
    const int rc = launch_stats_nt<128>(c, s, st, a, v, &c->stats_smem_set128, /*timed=*/false, /*max_grid=*/c->n_sms);
    TRY(tock(eb, st));
    return rc;
}

static int launch_stats(sqk_ctx *c, Slot &s, cudaStream_t st, const View &v, int mode, int lo, int hi, int num,
                        double std_scale, int32_t *d_nkept, const double *d_pa_off = nullptr,
                        const double *d_pa_scale = nullptr, int t_start = 0, int t_end = 0, StatsOut *so = nullptr,
                        bool want_mask = false)
{
    lo = clamp_lim(lo); hi = clamp_lim(hi);
    TRY(ensure(s.stats, (size_t)v.n_reads * sizeof(ReadStats)));
    StatsArgs a{};
    a.base = v.base; a.alloc_lo = v.alloc_lo; a.alloc_hi = v.alloc_hi;
    a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = v.n_reads;
    a.stats = (ReadStats *)s.stats.p; a.n_kept_out = d_nkept;
    a.mode = mode; a.lo = lo; a.hi = hi; a.num = num; a.std_scale = std_scale;
    a.pa_offset = d_pa_off; a.pa_scale = d_pa_scale;
    a.t_start = t_start; a.t_end = t_end;
    static int env_gen = -1;
    if (env_gen < 0) { const char *e = getenv("SQK_STATS_GEN"); env_gen = e ? atoi(e) : 0; }   // experiments: 1 = first generation only
    const int gen = c->stats_gen ? c->stats_gen : env_gen;
    StatsOut local;
    if (gen != 1 && mode != SQK_STATS_ADAPTER && v.max_len <= SQK_S2_MAX_LEN && v.n_reads < 0x7ffffff0LL) {
        // third generation (one warp per read) where it applies: raw-integer zscale / segmenter / none, histogram-sized window
        const bool s3_mode = mode == SQK_STATS_ZSCALE || mode == SQK_STATS_NONE || mode == SQK_STATS_MEDMAD || (mode == SQK_STATS_SEGMENTER && d_pa_off == nullptr);
        const int64_t span = (int64_t)std::min(hi - 1, 32767) - std::max(lo + 1, -32768) + 1;
        if (gen != 2 && s3_mode && ((mode != SQK_STATS_SEGMENTER && mode != SQK_STATS_MEDMAD) || span <= SQK_S3_MAX_BINS))
            return launch_stats3(c, s, st, a, v, want_mask, so ? so : &local);
        return launch_stats2(c, s, st, a, v, want_mask, so ? so : &local);
    }
    // one CTA per read; the warp-per-read form (SQK_STATS_NT=32, reads <= 8192 samples) is kept for experiments
    static int force_nt = -1;
    if (force_nt < 0) { const char *e = getenv("SQK_STATS_NT"); force_nt = e ? atoi(e) : 0; }
    const bool warp_per_read = force_nt == 32 && v.max_len <= SQK_HEAP_MAX_N;
    if (warp_per_read) return launch_stats_nt<32>(c, s, st, a, v, &c->stats_smem_set32);
    // Reads so long that their shared-memory staging (2 bytes/sample) leaves room for fewer than four 128-thread CTAs
    // per SM get 256 threads each: the same staged reads per SM, twice the warps working on them.
    const int64_t stage_bytes = 2 * std::min<int64_t>(v.max_len, ((int64_t)c->smem_optin - 4096) / 2) + 4096;
    const bool wide = force_nt == 256 || (force_nt == 0 && 4 * stage_bytes > (int64_t)c->smem_optin);
    return wide ? launch_stats_nt<256>(c, s, st, a, v, &c->stats_smem_set256)
                : launch_stats_nt<128>(c, s, st, a, v, &c->stats_smem_set128);
}

static int pick_dtw(const sqk_ctx *c, int N, int precision, int *L_out, int *K_out, sqk_dtw_launcher *fn)
{
    static const int kmin[5] = {SQK_DTW_L1_KMIN, SQK_DTW_L4_KMIN, SQK_DTW_L8_KMIN, SQK_DTW_L16_KMIN, SQK_DTW_L32_KMIN};
    static const int kmax[5] = {SQK_DTW_L1_KMAX, SQK_DTW_L4_KMAX, SQK_DTW_L8_KMAX, SQK_DTW_L16_KMAX, SQK_DTW_L32_KMAX};
    static const int lanes[5] = {1, 4, 8, 16, 32};
    static const sqk_dtw_launcher f64[5] = {sqk_launch_dtw_f64_l1, sqk_launch_dtw_f64_l4, sqk_launch_dtw_f64_l8,
                                            sqk_launch_dtw_f64_l16, sqk_launch_dtw_f64_l32};
    auto fits = [&](int li) {
        const int L = lanes[li], K = (N + L - 1) / L;
        if (K < kmin[li] || K > kmax[li]) return false;
        if (L * K - N > 0 && K < 2) return false;
        return true;
    };
    int pick = -1;
    if (c->force_lanes) {
        for (int li = 0; li < 5; li++) if (lanes[li] == c->force_lanes && fits(li)) pick = li;
        if (pick < 0) return fail(SQK_ERR_UNSUPPORTED, "motif of %d points cannot run with %d lanes per read", N, c->force_lanes);
    } else {
        // default policy (measured on B200, profiles/): 10-12 motif rows per lane keep the kernel at
        // <= 110 registers (4+ CTAs per SM) while one shuffle still pays for 10 cells; beyond 320 points
        // every lane is in use and the rows per lane grow instead.
        const int want_k = 12;
        for (int li = 0; li < 5 && pick < 0; li++)
            if (fits(li) && (N + lanes[li] - 1) / lanes[li] <= want_k) pick = li;
        for (int li = 0; li < 5 && pick < 0; li++)
            if (fits(li)) pick = li;
        if (pick < 0) return fail(SQK_ERR_UNSUPPORTED, "motif of %d points is longer than this build supports (max %d)", N,
                                  32 * SQK_DTW_L32_KMAX);
    }
    *L_out = lanes[pick];
    *K_out = (N + lanes[pick] - 1) / lanes[pick];
    (void)precision;      // SQK_PREC_FP32 is served by the exact path (see include/sqk.h)
    *fn = f64[pick];
    return SQK_OK;
}

// Lanes / rows-per-lane of the lower-bound kernel: same policy as the float64 kernel (fewest lanes with <= 12 rows per
// lane); SQK_LB_LANES overrides it for experiments.
static bool pick_lb_shape(const sqk_ctx *c, int N, int *L_out, int *K_out)
{
    static const int kmin[4] = {SQK_DTW_L4_KMIN, SQK_DTW_L8_KMIN, SQK_DTW_L16_KMIN, SQK_DTW_L32_KMIN};
    static const int kmax[4] = {SQK_DTW_L4_KMAX, SQK_DTW_L8_KMAX, SQK_DTW_L16_KMAX, SQK_DTW_L32_KMAX};
    static const int lanes[4] = {4, 8, 16, 32};
    static int env_lanes = -1;
    if (env_lanes < 0) { const char *e = getenv("SQK_LB_LANES"); env_lanes = e ? atoi(e) : 0; }   // experiments
    const int force = env_lanes ? env_lanes : c->force_lanes;
    auto fits = [&](int li) {
        const int L = lanes[li], K = (N + L - 1) / L;
        return K >= kmin[li] && K <= kmax[li] && K >= 2;
    };
    int pick = -1;
    if (force) {
        for (int li = 0; li < 4; li++) if (lanes[li] == force && fits(li)) pick = li;
    }
    if (pick < 0) {
        const int want_k = c->lb_want_k;
        for (int li = 0; li < 4 && pick < 0; li++)
            if (fits(li) && (N + lanes[li] - 1) / lanes[li] <= want_k) pick = li;
        for (int li = 0; li < 4 && pick < 0; li++)
            if (fits(li)) pick = li;
    }
    if (pick < 0) return false;
    *L_out = lanes[pick];
    *K_out = (N + lanes[pick] - 1) / lanes[pick];
    return true;
}

static sqk_lb_launcher pick_lb(int L)
{
    switch (L) {
    case 4: return sqk_launch_lb_l4;
    case 8: return sqk_launch_lb_l8;
    case 16: return sqk_launch_lb_l16;
    default: return sqk_launch_lb_l32;
    }
}

// float64 launcher with the most lanes per read that the motif length allows (at least 2 rows per lane)
static void pick_dtw_wide(int N, int *L_io, int *K_io, sqk_dtw_launcher *fn_io)
{
    static const int kmin[3] = {SQK_DTW_L8_KMIN, SQK_DTW_L16_KMIN, SQK_DTW_L32_KMIN};
    static const int kmax[3] = {SQK_DTW_L8_KMAX, SQK_DTW_L16_KMAX, SQK_DTW_L32_KMAX};
    static const int lanes[3] = {8, 16, 32};
    static const sqk_dtw_launcher f64[3] = {sqk_launch_dtw_f64_l8, sqk_launch_dtw_f64_l16, sqk_launch_dtw_f64_l32};
    for (int li = 2; li >= 0; li--) {
        const int L = lanes[li], K = (N + L - 1) / L;
        if (L <= *L_io) return;                       // not wider than the default
        if (K < kmin[li] || K > kmax[li] || K < 2) continue;
        *L_io = L; *K_io = K; *fn_io = f64[li];
        return;
    }
}

// Exact (float64) requests run as the two-pass plan when the reads are long enough for windows to pay:
// the float32 lower-bound scan + float64 windows (sqk_dtw_plan.cuh).  Same results bit for bit.
static bool want_two_pass(const sqk_ctx *c, const sqk_motif_params *p, int N, int64_t max_len)
{
    if (c->dtw_plan == SQK_PLAN_SINGLE_PASS) return false;
    if (c->dtw_plan == SQK_PLAN_TWO_PASS) return true;
    return max_len >= 4ll * (sqk_lb_window(N) + N);
}

static int check_motif_params(const sqk_motif_params *p)
{
    if (!p) return fail(SQK_ERR_ARG, "params is NULL");
    if (p->scale_mode < 0 || p->scale_mode > 2) return fail(SQK_ERR_ARG, "scale_mode %d not in {0,1,2}", p->scale_mode);
    if (p->precision < 0 || p->precision > 1) return fail(SQK_ERR_ARG, "precision %d not in {0,1}", p->precision);
    return SQK_OK;
}

#define SQK_CTRS_PER_MODEL 8   // [0] read queue head, [1] #window jobs, [2] window queue head, [3] #fallback jobs, [4] fallback queue head, [5] #second-attempt jobs, [6] their queue head, [7] row-block queue head

// A motif of more than 1024 points: its rows are cut into nb blocks of <= 1024; block b runs as one launch of the float64
// kernel's row-block variant over a sub-batch of reads, reading row r0 - 1 of every column from the boundary buffer the
// previous launch wrote (the free-start row for b = 0) and writing its own last row (the hit for b = nb - 1).  Same
// recurrence, same tie-breaks: the result is mlpy's bit for bit.  Two boundary buffers of 16 bytes per column per read.
static int enqueue_long_motif(sqk_ctx *c, Slot &s, cudaStream_t st, const View &v, const double *d_model, int N,
                              const sqk_motif_params *p, sqk_hit *d_hits, int hit_stride, unsigned *d_counter)
{
    const int nb = (N + 32 * SQK_DTW_L32_KMAX - 1) / (32 * SQK_DTW_L32_KMAX);
    const int rows = (N + nb - 1) / nb;
    const int64_t stride = (std::max<int64_t>(v.max_len, 1) + 7) & ~7ll;
    int64_t nsub = std::max<int64_t>(1, (1ll << 30) / (16 * stride));      // <= 1 GiB per boundary buffer
    nsub = std::min<int64_t>(nsub, v.n_reads);
    TRY(ensure(s.bnd_a, (size_t)nsub * stride * sizeof(BndCell)));
    TRY(ensure(s.bnd_b, (size_t)nsub * stride * sizeof(BndCell)));
    for (int64_t r0 = 0; r0 < v.n_reads; r0 += nsub) {
        const int64_t nr = std::min<int64_t>(nsub, v.n_reads - r0);
        BndCell *cur = (BndCell *)s.bnd_a.p, *nxt = (BndCell *)s.bnd_b.p;
        for (int b = 0; b < nb; b++) {
            DtwArgs a{};
            a.base = v.base; a.alloc_lo = v.alloc_lo; a.alloc_hi = v.alloc_hi;
            a.offsets = v.offsets; a.read0 = v.read0 + r0; a.n_reads = (int)nr;
            a.stats = (const ReadStats *)s.stats.p + r0;
            a.model = d_model + (size_t)b * rows; a.N = std::min(rows, N - b * rows);
            a.lo = clamp_lim(p->lo); a.hi = clamp_lim(p->hi);
            a.hits = d_hits + r0 * hit_stride; a.hit_stride = hit_stride;
            a.counter = d_counter;
            a.bnd_in = b > 0 ? cur : nullptr;
            a.bnd_out = b < nb - 1 ? nxt : nullptr;
            a.bnd_stride = stride;
            CU(cudaMemsetAsync(d_counter, 0, sizeof(unsigned), st));
            const int K = (a.N + 31) / 32;
            cudaError_t e = sqk_launch_dtw_f64_bnd_l32(K, a, c->n_sms, st);
            if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW row-block launch (N=%d, block %d of %d, K=%d): %s", N, b, nb, K, cudaGetErrorString(e));
            c->n_launches++;
            std::swap(cur, nxt);
        }
    }
    return SQK_OK;
}

// stats + the DTW of every model over a device-resident View; d_hits is [n_reads][n_models]
static int enqueue_motifseq(sqk_ctx *c, Slot &s, cudaStream_t st, const View &v, const double *d_models,
                            const double *h_models, const int32_t *h_model_offsets, int n_models,
                            const sqk_motif_params *p, sqk_hit *d_hits, int32_t *d_nkept, bool publish = false)
{
    if (v.n_reads == 0) return SQK_OK;
    if (v.n_reads > 0x7fffffffLL) return fail(SQK_ERR_ARG, "more than 2^31-1 reads in one launch");
    TRY(launch_stats(c, s, st, v, p->scale_mode, p->lo, p->hi, 0, 0.0, d_nkept));
    if (n_models > 256) return fail(SQK_ERR_UNSUPPORTED, "more than 256 models per call");
    TRY(ensure(s.counter, 256 * SQK_CTRS_PER_MODEL * sizeof(unsigned)));
    CU(cudaMemsetAsync(s.counter.p, 0, (size_t)n_models * SQK_CTRS_PER_MODEL * sizeof(unsigned), st));
    for (int m = 0; m < n_models; m++) {
        const int N = h_model_offsets[m + 1] - h_model_offsets[m];
        int L = 0, K = 0;
        sqk_dtw_launcher fn = nullptr;
        if (N <= 32 * SQK_DTW_L32_KMAX) TRY(pick_dtw(c, N, p->precision, &L, &K, &fn));
        unsigned *ctr = (unsigned *)s.counter.p + (size_t)m * SQK_CTRS_PER_MODEL;
        if (N > 32 * SQK_DTW_L32_KMAX) {
            // ---- motif longer than one pass holds (mlpy takes any length; MotifSeq.py:382-405 expands a fasta to ~9 points
            // per base): float64 row blocks, the last row of a block handed to the next one through a boundary row ---------
            cudaEvent_t eb;
            TRY(tick(c, SQK_K_DTW, st, &eb));
            TRY(enqueue_long_motif(c, s, st, v, d_models + h_model_offsets[m], N, p, d_hits + m, n_models, ctr + 7));
            if (publish && c->peers.n) {
                PeerOut po{};
                for (int q = 0; q < c->peers.n; q++) po.peer[po.n++] = c->peers.peer[q] + c->peer_base * n_models + m;
                const unsigned pg = (unsigned)std::max<int64_t>(1, std::min<int64_t>((v.n_reads + 255) / 256, 2 * c->n_sms));
                sqk_publish_kernel<<<pg, 256, 0, st>>>(d_hits + m, n_models, nullptr, nullptr, (int)v.n_reads, po);
                CU(cudaGetLastError());
                c->n_launches++;
            }
            TRY(tock(eb, st));
            continue;
        }
        DtwArgs a{};
        a.base = v.base; a.alloc_lo = v.alloc_lo; a.alloc_hi = v.alloc_hi;
        a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = (int)v.n_reads;
        a.stats = (const ReadStats *)s.stats.p;
        a.model = d_models + h_model_offsets[m]; a.N = N;
        a.lo = clamp_lim(p->lo); a.hi = clamp_lim(p->hi);
        a.hits = d_hits + m; a.hit_stride = n_models;
        a.counter = ctr;
        PeerOut po{};         // this model's column of this rank's block in every peer's gathered buffer
        if (publish)
            for (int q = 0; q < c->peers.n; q++) po.peer[po.n++] = c->peers.peer[q] + c->peer_base * n_models + m;
        const unsigned pub_grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((v.n_reads + 255) / 256, 2 * c->n_sms));
        cudaEvent_t eb;
        int LL = 0, LK = 0;   // shape of the lower-bound kernel
        if (L < 4 || !want_two_pass(c, p, N, v.max_len) || !pick_lb_shape(c, N, &LL, &LK)) {   // motifs of <= 4 points run one thread per read: single pass
            TRY(tick(c, SQK_K_DTW, st, &eb));
            cudaError_t e = fn(K, a, c->n_sms, st);
            if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW launch (N=%d, K=%d, L=%d): %s", N, K, L, cudaGetErrorString(e));
            c->n_launches++;
            if (po.n) {
                sqk_publish_kernel<<<pub_grid, 256, 0, st>>>(a.hits, a.hit_stride, nullptr, nullptr, (int)v.n_reads, po);
                CU(cudaGetLastError());
                c->n_launches++;
            }
            TRY(tock(eb, st));
            continue;
        }
        // ---- two-pass plan: lower-bound scan -> exact windows -> finalize -> exact fallback ----------------
        if ((int64_t)v.n_reads * SQK_LB_MAX_CLUSTERS > 0x7fffffffLL) return fail(SQK_ERR_ARG, "too many reads in one launch for the two-pass plan");
        TRY(ensure(s.jobs, (size_t)v.n_reads * SQK_LB_MAX_CLUSTERS * sizeof(DtwJob)));
        TRY(ensure(s.fbjobs, (size_t)v.n_reads * sizeof(DtwJob)));
        TRY(ensure(s.lbreads, (size_t)v.n_reads * sizeof(LbRead)));
        TRY(ensure(s.jobres, (size_t)v.n_reads * SQK_LB_MAX_CLUSTERS * sizeof(sqk_hit)));
        TRY(ensure(s.jobs2, (size_t)v.n_reads * SQK_LB_MAX_CLUSTERS * sizeof(DtwJob)));
        TRY(ensure(s.rtjobs, (size_t)v.n_reads * SQK_LB_MAX_CLUSTERS * sizeof(DtwJob)));
        TRY(ensure(s.pending, (size_t)v.n_reads));
        double xmax = 0.0;
        for (int i = h_model_offsets[m]; i < h_model_offsets[m + 1]; i++) xmax = std::max(xmax, std::fabs(h_models[i]));
        if (!(xmax < 1e30)) return fail(SQK_ERR_ARG, "model %d holds a non-finite point", m);
        LbArgs b{};
        b.base = v.base; b.alloc_lo = v.alloc_lo; b.alloc_hi = v.alloc_hi;
        b.offsets = v.offsets; b.read0 = v.read0; b.n_reads = (int)v.n_reads;
        b.stats = a.stats; b.model = a.model; b.N = N; b.lo = a.lo; b.hi = a.hi;
        b.hits = a.hits; b.hit_stride = a.hit_stride;
        b.counter = ctr;
        b.jobs = (DtwJob *)s.jobs.p; b.n_jobs = ctr + 1;
        b.reads = (LbRead *)s.lbreads.p;
        b.xmax_abs = xmax;
        b.W = sqk_lb_window(N);
        b.W2 = sqk_lb_window_retry(N);
        if (const char *e = getenv("SQK_LB_WINDOW")) { const int wv = atoi(e); if (wv > 0) b.W = wv; }   // test knob: small windows force the second attempt / fallback
        if (const char *e = getenv("SQK_LB_WINDOW2")) { const int wv = atoi(e); b.W2 = wv > 0 ? wv : 0; } // test knob: 0 = no second attempt
        b.jobs2 = b.W2 > 0 ? (DtwJob *)s.jobs2.p : nullptr;
        b.short_len = 2 * (b.W + N);
        { static int env_cols = -1; if (env_cols < 0) { const char *ec = getenv("SQK_LB_COLS"); env_cols = ec ? atoi(ec) : 0; } b.cols = (env_cols == 2 || env_cols == 4) ? env_cols : 0; }   // experiments
        TRY(tick(c, SQK_K_DTW_LB, st, &eb));
        cudaError_t e = pick_lb(LL)(LK, b, c->n_sms, st);
        if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW lower-bound launch (N=%d, K=%d, L=%d): %s", N, LK, LL, cudaGetErrorString(e));
        TRY(tock(eb, st));

        TRY(tick(c, SQK_K_DTW_WIN, st, &eb));
        a.counter = ctr + 2;
        a.jobs = b.jobs; a.n_jobs = ctr + 1;
        a.job_out = (sqk_hit *)s.jobres.p; a.job_out_stride = 1;
        e = fn(K, a, c->n_sms, st);
        if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW window launch (N=%d, K=%d, L=%d): %s", N, K, L, cudaGetErrorString(e));
        // wide-lane launcher for the (few) second-attempt and full-length jobs: they run at the latency of one job, so each
        // gets as many lanes as the motif allows (N = 80: 3 rows on 32 lanes instead of 10 on 8 -- same bits, a third of the time)
        int FL = L, FK = K;
        sqk_dtw_launcher ffn = fn;
        if (!c->force_lanes) pick_dtw_wide(N, &FL, &FK, &ffn);
        FinalizeArgs f{};
        f.reads = b.reads; f.jobres = (const sqk_hit *)s.jobres.p; f.n_reads = (int)v.n_reads;
        f.hits = a.hits; f.hit_stride = a.hit_stride;
        f.base = v.base; f.offsets = v.offsets; f.read0 = v.read0; f.stats = a.stats;
        f.fb_jobs = (DtwJob *)s.fbjobs.p; f.n_fb = ctr + 3;
        f.stage = 1; f.jobs2 = b.jobs2; f.rt_jobs = (DtwJob *)s.rtjobs.p; f.n_rt = ctr + 5; f.pending = (unsigned char *)s.pending.p;
        f.po = po;
        const unsigned fgrid = (unsigned)((v.n_reads + 255) / 256);
        sqk_dtw_finalize_kernel<<<fgrid, 256, 0, st>>>(f);
        CU(cudaGetLastError());
        if (b.jobs2) {
            // second attempt: the reads whose windows tainted, behind windows of W2 columns; then decide again
            a.counter = ctr + 6;
            a.jobs = f.rt_jobs; a.n_jobs = ctr + 5;
            a.job_out = (sqk_hit *)s.jobres.p; a.job_out_stride = 1;
            e = ffn(FK, a, c->n_sms, st);
            if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW second-attempt launch (N=%d, K=%d, L=%d): %s", N, FK, FL, cudaGetErrorString(e));
            f.stage = 2;
            sqk_dtw_finalize_kernel<<<fgrid, 256, 0, st>>>(f);
            CU(cudaGetLastError());
            c->n_launches += 2;
        }
        a.counter = ctr + 4;
        a.jobs = f.fb_jobs; a.n_jobs = ctr + 3;
        a.job_out = a.hits; a.job_out_stride = a.hit_stride;
        e = ffn(FK, a, c->n_sms, st);
        if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW fallback launch (N=%d, K=%d, L=%d): %s", N, FK, FL, cudaGetErrorString(e));
        c->n_launches += 4;                      // lower-bound scan, windows, finalize, fallback
        if (po.n) {                              // the (few) reads the fallback just wrote
            sqk_publish_kernel<<<8, 256, 0, st>>>(a.hits, a.hit_stride, f.fb_jobs, ctr + 3, (int)v.n_reads, po);
            CU(cudaGetLastError());
            c->n_launches++;
        }
        TRY(tock(eb, st));
    }
    return SQK_OK;
}

static int check_seg_params(const sqk_seg_params *p)
{
    if (!p) return fail(SQK_ERR_ARG, "params is NULL");
    if (p->max_segs < 1) return fail(SQK_ERR_ARG, "max_segs must be >= 1");
    if (p->corrector < 0) return fail(SQK_ERR_ARG, "corrector must be >= 0");
    if (p->window < 0 || p->error < 0) return fail(SQK_ERR_ARG, "window and error must be >= 0");
    return SQK_OK;
}

static int enqueue_segmenter(sqk_ctx *c, Slot &s, cudaStream_t st, const View &v, const sqk_seg_params *p,
                             int32_t *d_segs, int32_t *d_nsegs, const double *d_pa_off, const double *d_pa_scale)
{
    if (v.n_reads == 0) return SQK_OK;
    if (v.n_reads > 0x7fffffffLL) return fail(SQK_ERR_ARG, "more than 2^31-1 reads in one launch");
    const double fm = std::ceil((double)p->window * p->stall_len);
    const int first_min = fm > 2e9 ? 2000000000 : (fm < -2e9 ? -2000000000 : (int)fm);
    StatsOut so;
    TRY(launch_stats(c, s, st, v, SQK_STATS_SEGMENTER, p->lim_lo, p->lim_hi, p->num, p->std_scale, nullptr, d_pa_off,
                     d_pa_scale, 0, 0, &so, /*want_mask=*/true));
    FsmArgs a{};
    a.base = v.base; a.alloc_lo = v.alloc_lo; a.alloc_hi = v.alloc_hi;
    a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = (int)v.n_reads;
    a.stats = (const ReadStats *)s.stats.p;
    a.num = p->num;
    a.error = p->error; a.corrector = p->corrector; a.window = p->window; a.seg_dist = p->seg_dist;
    a.first_min = first_min;
    a.max_segs = p->max_segs; a.segs = d_segs; a.n_segs = d_nsegs;
    const unsigned grid = (unsigned)((v.n_reads + SQK_FSM_THREADS - 1) / SQK_FSM_THREADS);
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_SEG_FSM, st, &eb));
    if (so.v2 && so.mask) {
        // the state machine on the bit masks sqk_stats2_kernel emitted; the (few) reads of the redo list sample by sample
        FsmMaskArgs m{};
        m.mask = so.mask; m.mask_stride = so.mask_stride; m.n_reads = (int)v.n_reads; m.stats = a.stats;
        m.error = a.error; m.corrector = a.corrector; m.window = a.window; m.seg_dist = a.seg_dist;
        m.first_min = a.first_min; m.max_segs = a.max_segs; m.segs = d_segs; m.n_segs = d_nsegs;
        sqk_fsm_mask_kernel<<<grid, SQK_FSM_THREADS, 0, st>>>(m);
        CU(cudaGetLastError());
        c->n_launches++;
        a.list = so.redo; a.n_list = so.n_redo;
    }
    sqk_fsm_kernel<<<grid, SQK_FSM_THREADS, 0, st>>>(a);
    CU(cudaGetLastError());
    c->n_launches++;
    TRY(tock(eb, st));
    return SQK_OK;
}

static int check_adapter_params(const sqk_adapter_params *p)
{
    if (!p) return fail(SQK_ERR_ARG, "params is NULL");
    if (p->corrector < 1) return fail(SQK_ERR_ARG, "corrector must be >= 1");
    if (p->window < 0 || p->error < 0) return fail(SQK_ERR_ARG, "window and error must be >= 0");
    if (p->t_start < 0 || p->t_end < 0) return fail(SQK_ERR_ARG, "t_start and t_end must be >= 0");
    return SQK_OK;
}

// dRNA adapter finder: statistics of kept samples [t_start, t_end) (K1, SQK_STATS_ADAPTER) + its state machine
static int enqueue_adapter(sqk_ctx *c, Slot &s, cudaStream_t st, const View &v, const sqk_adapter_params *p,
                           int32_t *d_segs, int32_t *d_found)
{
    if (v.n_reads == 0) return SQK_OK;
    if (v.n_reads > 0x7fffffffLL) return fail(SQK_ERR_ARG, "more than 2^31-1 reads in one launch");
    TRY(launch_stats(c, s, st, v, SQK_STATS_ADAPTER, p->lim_lo, p->lim_hi, 0, p->std_scale, nullptr, nullptr, nullptr,
                     p->t_start, p->t_end));
    AdapterArgs a{};
    a.base = v.base; a.alloc_lo = v.alloc_lo; a.alloc_hi = v.alloc_hi;
    a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = (int)v.n_reads;
    a.stats = (const ReadStats *)s.stats.p;
    a.error = p->error; a.no_err_thresh = p->no_err_thresh; a.corrector = p->corrector; a.window = p->window;
    a.seg_dist = p->seg_dist;
    a.segs = d_segs; a.found = d_found;
    const unsigned grid = (unsigned)((v.n_reads + SQK_ADAPTER_THREADS - 1) / SQK_ADAPTER_THREADS);
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_SEG_FSM, st, &eb));
    sqk_adapter_fsm_kernel<<<grid, SQK_ADAPTER_THREADS, 0, st>>>(a);
    CU(cudaGetLastError());
    c->n_launches++;
    TRY(tock(eb, st));
    return SQK_OK;
}

static int check_rollmean_params(const sqk_rollmean_params *p)
{
    if (!p) return fail(SQK_ERR_ARG, "params is NULL");
    if (p->w < 1 || p->w > 65536) return fail(SQK_ERR_ARG, "w must be in [1, 65536]");
    return SQK_OK;
}

// dRNA rolling-mean adapter finder (sqk_rollmean.cuh): one kernel, one CTA per read
static int enqueue_rollmean(sqk_ctx *c, Slot &s, cudaStream_t st, const View &v, const sqk_rollmean_params *p,
                            int32_t *d_segs, int32_t *d_found)
{
    if (v.n_reads == 0) return SQK_OK;
    if (v.n_reads > 0x7fffffffLL) return fail(SQK_ERR_ARG, "more than 2^31-1 reads in one launch");
    if (v.max_len > 0x7ffffff0LL) return fail(SQK_ERR_UNSUPPORTED, "a read has more than 2^31-16 samples");
    const int dyn = (int)((sizeof(RmShared) + 15) & ~(size_t)15);
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sqk_rollmean_kernel, SQK_RM_THREADS, dyn));
    if (per_sm < 1) per_sm = 1;
    const int64_t p_stride = ((v.max_len + 1 + 31) & ~31ll), m_stride = ((v.max_len + 31) / 32 + 2 + 7) & ~7ll;
    // scratch per CTA: 4 bytes per sample (prefix sums) + two bit masks; keep the whole at <= 2 GiB
    const int64_t per_cta = 4 * p_stride + 8 * m_stride;
    int64_t grid = std::min<int64_t>(v.n_reads, (int64_t)c->n_sms * per_sm);
    grid = std::max<int64_t>(1, std::min<int64_t>(grid, (2ll << 30) / per_cta));
    TRY(ensure(s.rm_p, (size_t)grid * p_stride * sizeof(int32_t)));
    TRY(ensure(s.rm_masks, (size_t)grid * 2 * m_stride * sizeof(uint32_t)));
    RollmeanArgs a{};
    a.base = v.base; a.alloc_lo = v.alloc_lo; a.alloc_hi = v.alloc_hi;
    a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = (int)v.n_reads;
    a.lo = clamp_lim(p->lim_lo); a.hi = clamp_lim(p->lim_hi);
    a.w = p->w; a.seg_dist = p->seg_dist; a.lo_thresh = p->lo_thresh; a.hi_thresh = p->hi_thresh; a.shift = p->shift;
    a.std_factor = p->std_factor;
    a.P = (int32_t *)s.rm_p.p; a.p_stride = p_stride;
    a.masks = (uint32_t *)s.rm_masks.p; a.m_stride = m_stride;
    a.segs = d_segs; a.found = d_found;
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_SEG_FSM, st, &eb));
    sqk_rollmean_kernel<<<(unsigned)grid, SQK_RM_THREADS, dyn, st>>>(a);
    CU(cudaGetLastError());
    c->n_launches++;
    TRY(tock(eb, st));
    return SQK_OK;
}

// ---- float64-signal front end (sqk_f64.cuh): one extra kernel, then K2 / K3 as usual -------------------
struct View64 {
    const double *base;       // base[i] = absolute sample i
    const int64_t *offsets;   // indexable by absolute read id
    int64_t read0, n_reads;
    int64_t sample0, n_samples;   // the absolute sample range these reads cover
};

static int launch_f64_front(sqk_ctx *c, Slot &s, cudaStream_t st, const View64 &v, int mode, int lo, int hi, int num,
                            double std_scale, int32_t *d_nkept)
{
    TRY(ensure(s.stats, (size_t)v.n_reads * sizeof(ReadStats)));
    TRY(ensure(s.ynorm, (size_t)std::max<int64_t>(v.n_samples, 1) * sizeof(double)));
    if (mode == SQK_STATS_SEGMENTER) TRY(ensure(s.codes, (size_t)std::max<int64_t>(v.n_samples, 8) * sizeof(int16_t) + 32));
    F64Args a{};
    a.base = v.base; a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = v.n_reads;
    a.stats = (ReadStats *)s.stats.p; a.n_kept_out = d_nkept;
    a.mode = mode; a.lo = clamp_lim(lo); a.hi = clamp_lim(hi); a.num = num; a.std_scale = std_scale;
    a.ynorm = (double *)s.ynorm.p - v.sample0;
    a.codes = mode == SQK_STATS_SEGMENTER ? (int16_t *)s.codes.p - v.sample0 : nullptr;
    const int dyn = (int)((sizeof(StatsShared) + 15) & ~(size_t)15);
    const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(v.n_reads, (int64_t)c->n_sms * 8));
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_STATS, st, &eb));
    sqk_f64_front_kernel<<<(unsigned)grid, SQK_STATS_THREADS, dyn, st>>>(a);
    CU(cudaGetLastError());
    c->n_launches++;
    TRY(tock(eb, st));
    return SQK_OK;
}

static int enqueue_motifseq_f64(sqk_ctx *c, Slot &s, cudaStream_t st, const View64 &v, const double *d_models,
                                const int32_t *h_model_offsets, int n_models, const sqk_motif_params *p, sqk_hit *d_hits,
                                int32_t *d_nkept)
{
    if (v.n_reads == 0) return SQK_OK;
    if (v.n_reads > 0x7fffffffLL) return fail(SQK_ERR_ARG, "more than 2^31-1 reads in one launch");
    if (n_models > 256) return fail(SQK_ERR_UNSUPPORTED, "more than 256 models per call");
    TRY(launch_f64_front(c, s, st, v, p->scale_mode, p->lo, p->hi, 0, 0.0, d_nkept));
    TRY(ensure(s.counter, 256 * sizeof(unsigned)));
    CU(cudaMemsetAsync(s.counter.p, 0, 256 * sizeof(unsigned), st));
    for (int m = 0; m < n_models; m++) {
        const int N = h_model_offsets[m + 1] - h_model_offsets[m];
        int L = 0, K = 0;
        sqk_dtw_launcher fn = nullptr;
        TRY(pick_dtw(c, N, p->precision, &L, &K, &fn));
        DtwArgs a{};
        a.base = nullptr; a.alloc_lo = v.sample0; a.alloc_hi = v.sample0 + v.n_samples;
        a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = (int)v.n_reads;
        a.stats = (const ReadStats *)s.stats.p;
        a.model = d_models + h_model_offsets[m]; a.N = N;
        a.lo = 0; a.hi = 0;
        a.hits = d_hits + m; a.hit_stride = n_models;
        a.counter = (unsigned *)s.counter.p + m;
        a.prenorm = (const double *)s.ynorm.p - v.sample0;
        cudaEvent_t eb;
        TRY(tick(c, SQK_K_DTW, st, &eb));
        c->n_launches++;
        cudaError_t e = fn(K, a, c->n_sms, st);
        if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "DTW launch (N=%d, K=%d, L=%d): %s", N, K, L, cudaGetErrorString(e));
        TRY(tock(eb, st));
    }
    return SQK_OK;
}

static int enqueue_segmenter_f64(sqk_ctx *c, Slot &s, cudaStream_t st, const View64 &v, const sqk_seg_params *p,
                                 int32_t *d_segs, int32_t *d_nsegs)
{
    if (v.n_reads == 0) return SQK_OK;
    if (v.n_reads > 0x7fffffffLL) return fail(SQK_ERR_ARG, "more than 2^31-1 reads in one launch");
    TRY(launch_f64_front(c, s, st, v, SQK_STATS_SEGMENTER, p->lim_lo, p->lim_hi, p->num, p->std_scale, nullptr));
    const double fm = std::ceil((double)p->window * p->stall_len);
    FsmArgs a{};
    a.base = (const int16_t *)s.codes.p - v.sample0; a.alloc_lo = v.sample0; a.alloc_hi = v.sample0 + v.n_samples;
    a.offsets = v.offsets; a.read0 = v.read0; a.n_reads = (int)v.n_reads;
    a.stats = (const ReadStats *)s.stats.p;
    a.num = p->num; a.code_rows = 1;
    a.error = p->error; a.corrector = p->corrector; a.window = p->window; a.seg_dist = p->seg_dist;
    a.first_min = fm > 2e9 ? 2000000000 : (fm < -2e9 ? -2000000000 : (int)fm);
    a.max_segs = p->max_segs; a.segs = d_segs; a.n_segs = d_nsegs;
    const unsigned grid = (unsigned)((v.n_reads + SQK_FSM_THREADS - 1) / SQK_FSM_THREADS);
    cudaEvent_t eb;
    TRY(tick(c, SQK_K_SEG_FSM, st, &eb));
    sqk_fsm_kernel<<<grid, SQK_FSM_THREADS, 0, st>>>(a);
    CU(cudaGetLastError());
    TRY(tock(eb, st));
    return SQK_OK;
}

// first and last offset of a device-resident offsets array (one small synchronising copy)
static int device_sample_range(cudaStream_t st, const int64_t *d_offsets, int64_t n_reads, int64_t *s0, int64_t *s1)
{
    CU(cudaMemcpyAsync(s0, d_offsets, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(s1, d_offsets + n_reads, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (*s1 < *s0) return fail(SQK_ERR_ARG, "offsets are not non-decreasing");
    return SQK_OK;
}

static int device_max_len(sqk_ctx *c, cudaStream_t st, const int64_t *d_offsets, int64_t n_reads, int64_t *out)
{
    TRY(ensure(c->scratch8, 8));
    CU(cudaMemsetAsync(c->scratch8.p, 0, 8, st));
    const int grid = (int)std::min<int64_t>(4 * c->n_sms, (n_reads + 255) / 256);
    sqk_max_len_kernel<<<std::max(grid, 1), 256, 0, st>>>(d_offsets, n_reads, (unsigned long long *)c->scratch8.p);
    CU(cudaGetLastError());
    c->n_launches++;
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, c->scratch8.p, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (h == ~0ull) return fail(SQK_ERR_ARG, "offsets are not non-decreasing");
    *out = (int64_t)h;
    return SQK_OK;
}

// Split [0, n_reads) into chunks of about `target` samples (at least one read each).
static void plan_chunks(const int64_t *offsets, int64_t n_reads, int64_t target, std::vector<int64_t> &cuts)
{
    cuts.clear();
    cuts.push_back(0);
    int64_t r = 0;
    int ramp = 2;   // first chunks are 1/4 and 1/2 of the target: the GPU starts after a short first copy
    while (r < n_reads) {
        int64_t e = r + 1;
        const int64_t lim = offsets[r] + (target >> ramp);
        if (ramp > 0) ramp--;
        // gallop + binary search for the last read ending within the target
        int64_t lo = e, hi = n_reads;
        while (lo < hi) {
            const int64_t mid = lo + (hi - lo + 1) / 2;
            if (offsets[mid] <= lim) lo = mid; else hi = mid - 1;
        }
        e = std::max(e, lo);
        e = std::min(e, r + (int64_t)(1 << 22));   // bound reads per chunk (scratch sizes, int indices)
        cuts.push_back(e);
        r = e;
    }
}

static int64_t chunk_samples()   // samples per in-flight chunk; default 64 Mi (128 MiB of int16): >= one full wave of CTAs at 4k samples/read
{
    static int64_t v = 0;
    if (v == 0) {
        const char *e = getenv("SQK_CHUNK_SAMPLES");   // tuning knob for experiments
        v = e ? atoll(e) : 0;
        if (v < (1 << 16)) v = 64ll << 20;
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// exported API
// ------------------------------------------------------------------------------------------
extern "C" {

int sqk_version(void) { return SQK_VERSION; }

const char *sqk_last_error(void) { return g_err; }

int sqk_device_count(int *count)
{
    if (!count) return fail(SQK_ERR_ARG, "count is NULL");
    CU(cudaGetDeviceCount(count));
    return SQK_OK;
}

static int ctx_init(sqk_ctx *c, int device)
{
    c->device = device;
    int v = 0;
    CU(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); c->n_sms = v;
    CU(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); c->smem_optin = v;
    CU(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); c->clock_khz = v;
    CU(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device)); c->l2_bytes = v;
    int maj = 0, min = 0;
    CU(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, device));
    CU(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, device));
    c->cc = maj * 10 + min;
    if (c->cc != 100)
        return fail(SQK_ERR_UNSUPPORTED, "device %d is sm_%d; libsqk is built for sm_100a (B200) only", device, maj * 10 + min);
    for (int i = 0; i < 2; i++) CU(cudaStreamCreateWithFlags(&c->slot[i].stream, cudaStreamNonBlocking));
    return SQK_OK;
}

int sqk_ctx_create(int device, sqk_ctx **out)
{
    if (!out) return fail(SQK_ERR_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(SQK_ERR_CUDA, "no CUDA device available (%s); libsqk has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return fail(SQK_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    Guard g(device);
    if (!g.ok) return fail(SQK_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    sqk_ctx *c = new (std::nothrow) sqk_ctx();
    if (!c) return fail(SQK_ERR_NOMEM, "out of host memory");
    const int rc = ctx_init(c, device);
    if (rc != SQK_OK) {
        for (int i = 0; i < 2; i++) if (c->slot[i].stream) cudaStreamDestroy(c->slot[i].stream);
        delete c;
        return rc;
    }
    *out = c;
    return SQK_OK;
}

int sqk_ctx_destroy(sqk_ctx *c)
{
    if (!c) return SQK_OK;
    Guard g(c->device);
    cudaDeviceSynchronize();
    for (auto &r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto &e : c->pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) {
        Slot &s = c->slot[i];
        release(s.signals); release(s.offsets); release(s.stats); release(s.hits); release(s.nkept);
        release(s.segs); release(s.nsegs); release(s.counter); release(s.gstage); release(s.pa_off); release(s.pa_scale); release(s.ynorm); release(s.codes); release(s.rm_p); release(s.rm_masks); release(s.bnd_a); release(s.bnd_b); release(s.jobs2); release(s.rtjobs); release(s.pending);
        if (s.hin.p) { cudaFreeHost(s.hin.p); s.hin.p = nullptr; s.hin.cap = 0; }
        release(s.jobs); release(s.fbjobs); release(s.lbreads); release(s.jobres); release(s.redo); release(s.mask);
        for (int k = 0; k < 2; k++) if (s.hout[k].p) cudaFreeHost(s.hout[k].p);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    release(c->model); release(c->scratch8);
    if (c->model_stage.p) cudaFreeHost(c->model_stage.p);
    if (c->dev_done) cudaEventDestroy(c->dev_done);
    delete c;
    return SQK_OK;
}

int sqk_ctx_set_stream(sqk_ctx *c, void *cuda_stream)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    c->user_stream = (cudaStream_t)cuda_stream;
    c->use_user_stream = true;     // NULL is a valid stream (the legacy default stream, torch's default stream)
    return SQK_OK;
}

int sqk_ctx_reset_stream(sqk_ctx *c)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    c->user_stream = nullptr;
    c->use_user_stream = false;
    return SQK_OK;
}

static cudaStream_t device_stream(sqk_ctx *c) { return c->use_user_stream ? c->user_stream : c->slot[0].stream; }

int sqk_ctx_sync(sqk_ctx *c)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    Guard g(c->device);
    CU(cudaStreamSynchronize(device_stream(c)));
    CU(cudaStreamSynchronize(c->slot[0].stream));
    CU(cudaStreamSynchronize(c->slot[1].stream));
    return SQK_OK;
}

int sqk_ctx_device_props(sqk_ctx *c, int64_t props[5])
{
    if (!c || !props) return fail(SQK_ERR_ARG, "NULL argument");
    props[0] = c->n_sms; props[1] = c->smem_optin; props[2] = c->clock_khz; props[3] = c->l2_bytes; props[4] = c->cc;
    return SQK_OK;
}

int sqk_host_alloc(uint64_t bytes, void **out)
{
    if (!out) return fail(SQK_ERR_ARG, "out is NULL");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return SQK_OK;
}

int sqk_host_free(void *p)
{
    if (p) CU(cudaFreeHost(p));
    return SQK_OK;
}

int sqk_ctx_enable_timing(sqk_ctx *c, int on)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    c->timing = on != 0;
    return SQK_OK;
}

int sqk_ctx_get_timing(sqk_ctx *c, sqk_timing *out, int reset)
{
    if (!c || !out) return fail(SQK_ERR_ARG, "NULL argument");
    Guard g(c->device);
    TRY(drain_timing(c));
    *out = c->acc;
    if (reset) c->acc = sqk_timing{};
    return SQK_OK;
}

int sqk_ctx_set_chunk_samples(sqk_ctx *c, int64_t samples)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (samples < 0) return fail(SQK_ERR_ARG, "samples < 0");
    c->chunk_samples = samples;
    return SQK_OK;
}

int sqk_ctx_set_dtw_plan(sqk_ctx *c, int plan)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (plan != SQK_PLAN_AUTO && plan != SQK_PLAN_SINGLE_PASS && plan != SQK_PLAN_TWO_PASS)
        return fail(SQK_ERR_ARG, "plan must be SQK_PLAN_AUTO, SQK_PLAN_SINGLE_PASS or SQK_PLAN_TWO_PASS");
    c->dtw_plan = plan;
    return SQK_OK;
}

int sqk_ctx_get_plan_counters(sqk_ctx *c, int64_t out[2])
{
    if (!c || !out) return fail(SQK_ERR_ARG, "NULL argument");
    Guard g(c->device);
    out[0] = out[1] = 0;
    if (!c->slot[0].counter.p) return SQK_OK;
    CU(cudaStreamSynchronize(device_stream(c)));
    CU(cudaStreamSynchronize(c->slot[0].stream));
    unsigned h[SQK_CTRS_PER_MODEL];
    CU(cudaMemcpy(h, c->slot[0].counter.p, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h[1]; out[1] = h[3];
    return SQK_OK;
}

int sqk_ctx_get_plan_counters_ex(sqk_ctx *c, int64_t out[4])
{
    if (!c || !out) return fail(SQK_ERR_ARG, "NULL argument");
    Guard g(c->device);
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!c->slot[0].counter.p) return SQK_OK;
    CU(cudaStreamSynchronize(device_stream(c)));
    CU(cudaStreamSynchronize(c->slot[0].stream));
    unsigned h[SQK_CTRS_PER_MODEL];
    CU(cudaMemcpy(h, c->slot[0].counter.p, sizeof(h), cudaMemcpyDeviceToHost));
    out[0] = h[1]; out[1] = h[3]; out[2] = h[5];
    return SQK_OK;
}

int sqk_ctx_set_dtw_lanes(sqk_ctx *c, int lanes)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (lanes != 0 && lanes != 1 && lanes != 4 && lanes != 8 && lanes != 16 && lanes != 32)
        return fail(SQK_ERR_ARG, "lanes must be one of 0,1,4,8,16,32");
    c->force_lanes = lanes;
    return SQK_OK;
}

int sqk_ctx_get_launches(sqk_ctx *c, int64_t *out, int reset)
{
    if (!c || !out) return fail(SQK_ERR_ARG, "NULL argument");
    *out = c->n_launches;
    if (reset) c->n_launches = 0;
    return SQK_OK;
}

int sqk_ctx_set_stats_generation(sqk_ctx *c, int gen)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (gen < 0 || gen > 2) return fail(SQK_ERR_ARG, "generation must be 0 (automatic), 1 or 2");
    c->stats_gen = gen;
    return SQK_OK;
}

// ---- multi-GPU publication (one process per GPU) ------------------------------------------------------------------
int sqk_device_alloc(sqk_ctx *c, uint64_t bytes, void **out)
{
    if (!c || !out) return fail(SQK_ERR_ARG, "NULL argument");
    Guard g(c->device);
    *out = nullptr;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(SQK_ERR_NOMEM, "cudaMalloc(%llu bytes): %s", (unsigned long long)bytes, cudaGetErrorString(e)); }
    CU(cudaMemset(*out, 0, bytes ? bytes : 1));
    return SQK_OK;
}

int sqk_device_free(sqk_ctx *c, void *p)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    Guard g(c->device);
    if (p) CU(cudaFree(p));
    return SQK_OK;
}

int sqk_ipc_export(sqk_ctx *c, void *dev_ptr, unsigned char handle[64])
{
    if (!c || !dev_ptr || !handle) return fail(SQK_ERR_ARG, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    Guard g(c->device);
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle, &h, 64);
    return SQK_OK;
}

int sqk_ipc_open(sqk_ctx *c, const unsigned char handle[64], void **out)
{
    if (!c || !handle || !out) return fail(SQK_ERR_ARG, "NULL argument");
    Guard g(c->device);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return SQK_OK;
}

int sqk_ipc_close(sqk_ctx *c, void *p)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    Guard g(c->device);
    if (p) CU(cudaIpcCloseMemHandle(p));
    return SQK_OK;
}

int sqk_ctx_set_hit_peers_ex(sqk_ctx *c, void *const *peers, int n_peers, int64_t first_record, int64_t capacity_records)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (n_peers < 0 || n_peers > SQK_MAX_PEERS) return fail(SQK_ERR_ARG, "n_peers must be 0..%d", SQK_MAX_PEERS);
    if (n_peers > 0 && !peers) return fail(SQK_ERR_ARG, "peers is NULL");
    if (first_record < 0 || capacity_records < 0) return fail(SQK_ERR_ARG, "first_record / capacity_records < 0");
    c->peers = PeerOut{};
    for (int p = 0; p < n_peers; p++) {
        if (!peers[p]) return fail(SQK_ERR_ARG, "peers[%d] is NULL", p);
        c->peers.peer[p] = (sqk_hit *)peers[p];
    }
    c->peers.n = n_peers;
    c->peer_base = first_record;
    c->peer_cap = capacity_records;
    return SQK_OK;
}

int sqk_ctx_set_hit_peers(sqk_ctx *c, void *const *peers, int n_peers, int64_t first_record)
{
    return sqk_ctx_set_hit_peers_ex(c, peers, n_peers, first_record, 0);
}

int sqk_ctx_set_flag_peers(sqk_ctx *c, void *const *flag_arrays, int n_ranks, int my_rank)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (n_ranks < 0 || n_ranks > SQK_MAX_PEERS) return fail(SQK_ERR_ARG, "n_ranks must be 0..%d", SQK_MAX_PEERS);
    if (n_ranks > 0 && (!flag_arrays || my_rank < 0 || my_rank >= n_ranks)) return fail(SQK_ERR_ARG, "bad flag arguments");
    c->flags = PeerFlags{};
    for (int p = 0; p < n_ranks; p++) {
        if (!flag_arrays[p]) return fail(SQK_ERR_ARG, "flag_arrays[%d] is NULL", p);
        c->flags.arr[p] = (unsigned long long *)flag_arrays[p];
    }
    c->flags.n = n_ranks; c->flags.self = my_rank;
    return SQK_OK;
}

int sqk_peer_signal(sqk_ctx *c, uint64_t value)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (c->flags.n == 0) return SQK_OK;
    Guard g(c->device);
    cudaStream_t st = device_stream(c);
    TRY(dev_begin(c, st));
    sqk_peer_signal_kernel<<<1, 32, 0, st>>>(c->flags, (unsigned long long)value);
    CU(cudaGetLastError());
    c->n_launches++;
    TRY(dev_end(c, st));
    return SQK_OK;
}

int sqk_peer_wait(sqk_ctx *c, uint64_t value)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (c->flags.n == 0) return SQK_OK;
    Guard g(c->device);
    cudaStream_t st = device_stream(c);
    TRY(dev_begin(c, st));
    sqk_peer_wait_kernel<<<1, 32, 0, st>>>(c->flags, (unsigned long long)value);
    CU(cudaGetLastError());
    c->n_launches++;
    TRY(dev_end(c, st));
    return SQK_OK;
}

int sqk_motifseq(sqk_ctx *c, const int16_t *signals, const int64_t *offsets, int64_t n_reads, int64_t max_read_len,
                 const double *models, const int32_t *model_offsets, int32_t n_models, const sqk_motif_params *p, int mem,
                 sqk_hit *hits, int32_t *n_kept)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (n_reads < 0) return fail(SQK_ERR_ARG, "n_reads < 0");
    if (n_reads == 0) return SQK_OK;
    if (!offsets || !hits) return fail(SQK_ERR_ARG, "offsets/hits is NULL");
    if (!models || !model_offsets || n_models < 1) return fail(SQK_ERR_ARG, "no models");
    if (mem != SQK_MEM_HOST && mem != SQK_MEM_DEVICE) return fail(SQK_ERR_ARG, "mem must be SQK_MEM_HOST or SQK_MEM_DEVICE");
    TRY(check_motif_params(p));
    for (int m = 0; m < n_models; m++)
        if (model_offsets[m + 1] - model_offsets[m] < 1) return fail(SQK_ERR_ARG, "model %d is empty", m);
    Guard g(c->device);
    if (!g.ok) return fail(SQK_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);

    // models are tiny: staged from the host copy when they change (models / model_offsets are host pointers in both modes)
    const size_t model_points = (size_t)model_offsets[n_models];

    if (mem == SQK_MEM_DEVICE) {
        cudaStream_t st = device_stream(c);
        if (!signals) return fail(SQK_ERR_ARG, "signals is NULL");
        TRY(upload_models(c, models, model_points, st));
        TRY(dev_begin(c, st));
        if (max_read_len <= 0) TRY(device_max_len(c, st, offsets, n_reads, &max_read_len));
        View v{signals, 1, 0, offsets, 0, n_reads, max_read_len};   // bounds resolved on the device from offsets
        // device mode: the caller's allocation bounds are unknown, so [offsets[0], offsets[n]) delimits what the
        // kernels may touch (resolve_bounds): 16-byte blocks sticking out of it are read sample by sample.
        // Publication is armed for ONE call (sqk_ctx_set_hit_peers before every step): a later call that is not part of
        // the exchange -- a check against a single-GPU run, another batch shape -- must not write into the peers' buffers.
        const bool publish = c->peers.n > 0;
        if (publish && c->peer_cap > 0 && c->peer_base + n_reads > c->peer_cap) {
            c->peers.n = 0;
            return fail(SQK_ERR_ARG, "publication: records [%lld, %lld) do not fit the peers' buffers of %lld records",
                        (long long)c->peer_base, (long long)(c->peer_base + n_reads), (long long)c->peer_cap);
        }
        const int rc = enqueue_motifseq(c, c->slot[0], st, v, (const double *)c->model.p, models, model_offsets, n_models, p, hits, n_kept,
                                        publish);
        c->peers.n = 0;
        TRY(dev_end(c, st));
        return rc;
    }

    // ---- host mode: chunked, double-buffered H2D | stats+DTW | D2H ---------------------------
    if (!signals && offsets[n_reads] > offsets[0]) return fail(SQK_ERR_ARG, "signals is NULL");
    int64_t maxlen = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t d = offsets[r + 1] - offsets[r];
        if (d < 0) return fail(SQK_ERR_ARG, "offsets are not non-decreasing at read %lld", (long long)r);
        maxlen = std::max(maxlen, d);
    }
    if (maxlen > 0x7fffffffLL) return fail(SQK_ERR_UNSUPPORTED, "a read has more than 2^31-1 samples");
    TRY(host_begin(c));
    TRY(upload_models(c, models, model_points, c->slot[0].stream));
    std::vector<int64_t> cuts;
    plan_chunks(offsets, n_reads, c->chunk_samples > 0 ? c->chunk_samples : chunk_samples(), cuts);
    const bool hits_pinned = is_pinned(hits), nkept_pinned = n_kept && is_pinned(n_kept), sig_pinned = signals && is_pinned(signals);
    PipeGuard pg(c);
    for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
        Slot &s = c->slot[ci & 1];
        cudaStream_t st = s.stream;
        TRY(recycle(s));                 // the chunk that used this slot two iterations ago is done
        const int64_t r0 = cuts[ci], r1 = cuts[ci + 1], nr = r1 - r0;
        const int64_t s0 = offsets[r0], s1 = offsets[r1], ns = s1 - s0;
        TRY(ensure(s.signals, (size_t)std::max<int64_t>(ns, 8) * sizeof(int16_t) + 16));
        TRY(ensure(s.offsets, (size_t)(nr + 1) * sizeof(int64_t)));
        TRY(ensure(s.hits, (size_t)nr * n_models * sizeof(sqk_hit)));
        TRY(ensure(s.nkept, (size_t)nr * sizeof(int32_t)));
        TRY(signals_to_device(s, st, s.signals.p, signals + s0, (size_t)ns * sizeof(int16_t), sig_pinned));
        CU(cudaMemcpyAsync(s.offsets.p, offsets + r0, (size_t)(nr + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        View v{(const int16_t *)s.signals.p - s0, s0, s1, (const int64_t *)s.offsets.p - r0, r0, nr, maxlen};
        TRY(enqueue_motifseq(c, s, st, v, (const double *)c->model.p, models, model_offsets, n_models, p, (sqk_hit *)s.hits.p,
                             (int32_t *)s.nkept.p));
        TRY(result_to_host(s, 0, hits + r0 * n_models, s.hits.p, (size_t)nr * n_models * sizeof(sqk_hit), hits_pinned));
        if (n_kept) TRY(result_to_host(s, 1, n_kept + r0, s.nkept.p, (size_t)nr * sizeof(int32_t), nkept_pinned));
    }
    TRY(recycle(c->slot[0]));
    TRY(recycle(c->slot[1]));
    pg.armed = false;
    return SQK_OK;
}

}  // extern "C"

static int segmenter_impl(sqk_ctx *c, const int16_t *signals, const int64_t *offsets, int64_t n_reads, int64_t max_read_len,
                          const double *pa_offset, const double *pa_scale, const sqk_seg_params *p, int mem, int32_t *segs,
                          int32_t *n_segs, const sqk_adapter_params *ap = nullptr, const sqk_rollmean_params *rp = nullptr)
{
    if ((pa_offset == nullptr) != (pa_scale == nullptr)) return fail(SQK_ERR_ARG, "pa_offset and pa_scale must be given together");
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (n_reads < 0) return fail(SQK_ERR_ARG, "n_reads < 0");
    if (n_reads == 0) return SQK_OK;
    if (!offsets || !segs || !n_segs) return fail(SQK_ERR_ARG, "offsets/segs/n_segs is NULL");
    if (mem != SQK_MEM_HOST && mem != SQK_MEM_DEVICE) return fail(SQK_ERR_ARG, "mem must be SQK_MEM_HOST or SQK_MEM_DEVICE");
    if (ap) TRY(check_adapter_params(ap));      // adapter modes: one (start, end) row per read, n_segs = found flag
    else if (rp) TRY(check_rollmean_params(rp));
    else TRY(check_seg_params(p));
    Guard g(c->device);
    if (!g.ok) return fail(SQK_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
    const size_t seg_row = (size_t)((ap || rp) ? 1 : p->max_segs) * 2 * sizeof(int32_t);

    if (mem == SQK_MEM_DEVICE) {
        cudaStream_t st = device_stream(c);
        if (!signals) return fail(SQK_ERR_ARG, "signals is NULL");
        TRY(dev_begin(c, st));
        if (max_read_len <= 0) TRY(device_max_len(c, st, offsets, n_reads, &max_read_len));
        View v{signals, 1, 0, offsets, 0, n_reads, max_read_len};   // bounds resolved on the device from offsets
        int rc;
        if (ap) rc = enqueue_adapter(c, c->slot[0], st, v, ap, segs, n_segs);
        else if (rp) rc = enqueue_rollmean(c, c->slot[0], st, v, rp, segs, n_segs);
        else rc = enqueue_segmenter(c, c->slot[0], st, v, p, segs, n_segs, pa_offset, pa_scale);
        TRY(dev_end(c, st));
        return rc;
    }

    if (!signals && offsets[n_reads] > offsets[0]) return fail(SQK_ERR_ARG, "signals is NULL");
    int64_t maxlen = 0;
    for (int64_t r = 0; r < n_reads; r++) {
        const int64_t d = offsets[r + 1] - offsets[r];
        if (d < 0) return fail(SQK_ERR_ARG, "offsets are not non-decreasing at read %lld", (long long)r);
        maxlen = std::max(maxlen, d);
    }
    if (maxlen > 0x7fffffffLL) return fail(SQK_ERR_UNSUPPORTED, "a read has more than 2^31-1 samples");
    std::vector<int64_t> cuts;
    plan_chunks(offsets, n_reads, c->chunk_samples > 0 ? c->chunk_samples : chunk_samples(), cuts);
    const bool segs_pinned = is_pinned(segs), nsegs_pinned = is_pinned(n_segs), sig_pinned = signals && is_pinned(signals);
    TRY(host_begin(c));
    PipeGuard pg(c);
    for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
        Slot &s = c->slot[ci & 1];
        cudaStream_t st = s.stream;
        TRY(recycle(s));
        const int64_t r0 = cuts[ci], r1 = cuts[ci + 1], nr = r1 - r0;
        const int64_t s0 = offsets[r0], s1 = offsets[r1], ns = s1 - s0;
        TRY(ensure(s.signals, (size_t)std::max<int64_t>(ns, 8) * sizeof(int16_t) + 16));
        TRY(ensure(s.offsets, (size_t)(nr + 1) * sizeof(int64_t)));
        TRY(ensure(s.segs, (size_t)nr * seg_row));
        TRY(ensure(s.nsegs, (size_t)nr * sizeof(int32_t)));
        TRY(signals_to_device(s, st, s.signals.p, signals + s0, (size_t)ns * sizeof(int16_t), sig_pinned));
        CU(cudaMemcpyAsync(s.offsets.p, offsets + r0, (size_t)(nr + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(s.segs.p, 0, (size_t)nr * seg_row, st));
        const double *d_po = nullptr, *d_ps = nullptr;
        if (pa_offset) {
            TRY(ensure(s.pa_off, (size_t)nr * sizeof(double)));
            TRY(ensure(s.pa_scale, (size_t)nr * sizeof(double)));
            CU(cudaMemcpyAsync(s.pa_off.p, pa_offset + r0, (size_t)nr * sizeof(double), cudaMemcpyHostToDevice, st));
            CU(cudaMemcpyAsync(s.pa_scale.p, pa_scale + r0, (size_t)nr * sizeof(double), cudaMemcpyHostToDevice, st));
            d_po = (const double *)s.pa_off.p - r0; d_ps = (const double *)s.pa_scale.p - r0;
        }
        View v{(const int16_t *)s.signals.p - s0, s0, s1, (const int64_t *)s.offsets.p - r0, r0, nr, maxlen};
        if (ap) TRY(enqueue_adapter(c, s, st, v, ap, (int32_t *)s.segs.p, (int32_t *)s.nsegs.p));
        else if (rp) TRY(enqueue_rollmean(c, s, st, v, rp, (int32_t *)s.segs.p, (int32_t *)s.nsegs.p));
        else TRY(enqueue_segmenter(c, s, st, v, p, (int32_t *)s.segs.p, (int32_t *)s.nsegs.p, d_po, d_ps));
        TRY(result_to_host(s, 0, (char *)segs + (size_t)r0 * seg_row, s.segs.p, (size_t)nr * seg_row, segs_pinned));
        TRY(result_to_host(s, 1, n_segs + r0, s.nsegs.p, (size_t)nr * sizeof(int32_t), nsegs_pinned));
    }
    TRY(recycle(c->slot[0]));
    TRY(recycle(c->slot[1]));
    pg.armed = false;
    return SQK_OK;
}

extern "C" {

int sqk_motifseq_f64(sqk_ctx *c, const double *signals, const int64_t *offsets, int64_t n_reads, const double *models,
                     const int32_t *model_offsets, int32_t n_models, const sqk_motif_params *p, int mem, sqk_hit *hits,
                     int32_t *n_kept)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (n_reads < 0) return fail(SQK_ERR_ARG, "n_reads < 0");
    if (n_reads == 0) return SQK_OK;
    if (!offsets || !hits) return fail(SQK_ERR_ARG, "offsets/hits is NULL");
    if (!models || !model_offsets || n_models < 1) return fail(SQK_ERR_ARG, "no models");
    if (mem != SQK_MEM_HOST && mem != SQK_MEM_DEVICE) return fail(SQK_ERR_ARG, "mem must be SQK_MEM_HOST or SQK_MEM_DEVICE");
    TRY(check_motif_params(p));
    for (int m = 0; m < n_models; m++)
        if (model_offsets[m + 1] - model_offsets[m] < 1) return fail(SQK_ERR_ARG, "model %d is empty", m);
    Guard g(c->device);
    if (!g.ok) return fail(SQK_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
    const size_t model_points = (size_t)model_offsets[n_models];
    if (mem == SQK_MEM_DEVICE) {
        cudaStream_t st = device_stream(c);
        if (!signals) return fail(SQK_ERR_ARG, "signals is NULL");
        TRY(upload_models(c, models, model_points, st));
        TRY(dev_begin(c, st));
        int64_t s0 = 0, s1 = 0;
        TRY(device_sample_range(st, offsets, n_reads, &s0, &s1));
        View64 v{signals, offsets, 0, n_reads, s0, s1 - s0};
        const int rc = enqueue_motifseq_f64(c, c->slot[0], st, v, (const double *)c->model.p, model_offsets, n_models, p, hits, n_kept);
        TRY(dev_end(c, st));
        return rc;
    }
    if (!signals && offsets[n_reads] > offsets[0]) return fail(SQK_ERR_ARG, "signals is NULL");
    for (int64_t r = 0; r < n_reads; r++)
        if (offsets[r + 1] < offsets[r]) return fail(SQK_ERR_ARG, "offsets are not non-decreasing at read %lld", (long long)r);
    TRY(host_begin(c));
    TRY(upload_models(c, models, model_points, c->slot[0].stream));
    std::vector<int64_t> cuts;
    plan_chunks(offsets, n_reads, std::max<int64_t>((c->chunk_samples > 0 ? c->chunk_samples : chunk_samples()) / 4, 1), cuts);
    const bool hits_pinned = is_pinned(hits), nkept_pinned = n_kept && is_pinned(n_kept), sig_pinned = signals && is_pinned(signals);
    PipeGuard pg(c);
    for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
        Slot &s = c->slot[ci & 1];
        cudaStream_t st = s.stream;
        TRY(recycle(s));
        const int64_t r0 = cuts[ci], r1 = cuts[ci + 1], nr = r1 - r0;
        const int64_t s0 = offsets[r0], s1 = offsets[r1], ns = s1 - s0;
        TRY(ensure(s.signals, (size_t)std::max<int64_t>(ns, 2) * sizeof(double)));
        TRY(ensure(s.offsets, (size_t)(nr + 1) * sizeof(int64_t)));
        TRY(ensure(s.hits, (size_t)nr * n_models * sizeof(sqk_hit)));
        TRY(ensure(s.nkept, (size_t)nr * sizeof(int32_t)));
        TRY(signals_to_device(s, st, s.signals.p, signals + s0, (size_t)ns * sizeof(double), sig_pinned));
        CU(cudaMemcpyAsync(s.offsets.p, offsets + r0, (size_t)(nr + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        View64 v{(const double *)s.signals.p - s0, (const int64_t *)s.offsets.p - r0, r0, nr, s0, ns};
        TRY(enqueue_motifseq_f64(c, s, st, v, (const double *)c->model.p, model_offsets, n_models, p, (sqk_hit *)s.hits.p,
                                 (int32_t *)s.nkept.p));
        TRY(result_to_host(s, 0, hits + r0 * n_models, s.hits.p, (size_t)nr * n_models * sizeof(sqk_hit), hits_pinned));
        if (n_kept) TRY(result_to_host(s, 1, n_kept + r0, s.nkept.p, (size_t)nr * sizeof(int32_t), nkept_pinned));
    }
    TRY(recycle(c->slot[0]));
    TRY(recycle(c->slot[1]));
    pg.armed = false;
    return SQK_OK;
}

int sqk_segmenter_f64(sqk_ctx *c, const double *signals, const int64_t *offsets, int64_t n_reads,
                      const sqk_seg_params *p, int mem, int32_t *segs, int32_t *n_segs)
{
    if (!c) return fail(SQK_ERR_ARG, "ctx is NULL");
    if (n_reads < 0) return fail(SQK_ERR_ARG, "n_reads < 0");
    if (n_reads == 0) return SQK_OK;
    if (!offsets || !segs || !n_segs) return fail(SQK_ERR_ARG, "offsets/segs/n_segs is NULL");
    if (mem != SQK_MEM_HOST && mem != SQK_MEM_DEVICE) return fail(SQK_ERR_ARG, "mem must be SQK_MEM_HOST or SQK_MEM_DEVICE");
    TRY(check_seg_params(p));
    Guard g(c->device);
    if (!g.ok) return fail(SQK_ERR_CUDA, "cudaSetDevice(%d) failed", c->device);
    const size_t seg_row = (size_t)p->max_segs * 2 * sizeof(int32_t);
    if (mem == SQK_MEM_DEVICE) {
        cudaStream_t st = device_stream(c);
        if (!signals) return fail(SQK_ERR_ARG, "signals is NULL");
        TRY(dev_begin(c, st));
        int64_t s0 = 0, s1 = 0;
        TRY(device_sample_range(st, offsets, n_reads, &s0, &s1));
        View64 v{signals, offsets, 0, n_reads, s0, s1 - s0};
        const int rc = enqueue_segmenter_f64(c, c->slot[0], st, v, p, segs, n_segs);
        TRY(dev_end(c, st));
        return rc;
    }
    if (!signals && offsets[n_reads] > offsets[0]) return fail(SQK_ERR_ARG, "signals is NULL");
    for (int64_t r = 0; r < n_reads; r++)
        if (offsets[r + 1] < offsets[r]) return fail(SQK_ERR_ARG, "offsets are not non-decreasing at read %lld", (long long)r);
    std::vector<int64_t> cuts;
    plan_chunks(offsets, n_reads, std::max<int64_t>((c->chunk_samples > 0 ? c->chunk_samples : chunk_samples()) / 4, 1), cuts);
    const bool segs_pinned = is_pinned(segs), nsegs_pinned = is_pinned(n_segs), sig_pinned = signals && is_pinned(signals);
    TRY(host_begin(c));
    PipeGuard pg(c);
    for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
        Slot &s = c->slot[ci & 1];
        cudaStream_t st = s.stream;
        TRY(recycle(s));
        const int64_t r0 = cuts[ci], r1 = cuts[ci + 1], nr = r1 - r0;
        const int64_t s0 = offsets[r0], s1 = offsets[r1], ns = s1 - s0;
        TRY(ensure(s.signals, (size_t)std::max<int64_t>(ns, 2) * sizeof(double)));
        TRY(ensure(s.offsets, (size_t)(nr + 1) * sizeof(int64_t)));
        TRY(ensure(s.segs, (size_t)nr * seg_row));
        TRY(ensure(s.nsegs, (size_t)nr * sizeof(int32_t)));
        TRY(signals_to_device(s, st, s.signals.p, signals + s0, (size_t)ns * sizeof(double), sig_pinned));
        CU(cudaMemcpyAsync(s.offsets.p, offsets + r0, (size_t)(nr + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(s.segs.p, 0, (size_t)nr * seg_row, st));
        View64 v{(const double *)s.signals.p - s0, (const int64_t *)s.offsets.p - r0, r0, nr, s0, ns};
        TRY(enqueue_segmenter_f64(c, s, st, v, p, (int32_t *)s.segs.p, (int32_t *)s.nsegs.p));
        TRY(result_to_host(s, 0, (char *)segs + (size_t)r0 * seg_row, s.segs.p, (size_t)nr * seg_row, segs_pinned));
        TRY(result_to_host(s, 1, n_segs + r0, s.nsegs.p, (size_t)nr * sizeof(int32_t), nsegs_pinned));
    }
    TRY(recycle(c->slot[0]));
    TRY(recycle(c->slot[1]));
    pg.armed = false;
    return SQK_OK;
}

int sqk_segmenter(sqk_ctx *c, const int16_t *signals, const int64_t *offsets, int64_t n_reads, int64_t max_read_len,
                  const sqk_seg_params *p, int mem, int32_t *segs, int32_t *n_segs)
{
    return segmenter_impl(c, signals, offsets, n_reads, max_read_len, nullptr, nullptr, p, mem, segs, n_segs);
}

int sqk_adapter(sqk_ctx *c, const int16_t *signals, const int64_t *offsets, int64_t n_reads, int64_t max_read_len,
                const sqk_adapter_params *p, int mem, int32_t *segs, int32_t *found)
{
    if (!p) return fail(SQK_ERR_ARG, "params is NULL");
    return segmenter_impl(c, signals, offsets, n_reads, max_read_len, nullptr, nullptr, nullptr, mem, segs, found, p);
}

int sqk_rollmean(sqk_ctx *c, const int16_t *signals, const int64_t *offsets, int64_t n_reads, int64_t max_read_len,
                 const sqk_rollmean_params *p, int mem, int32_t *segs, int32_t *found)
{
    if (!p) return fail(SQK_ERR_ARG, "params is NULL");
    return segmenter_impl(c, signals, offsets, n_reads, max_read_len, nullptr, nullptr, nullptr, mem, segs, found, nullptr, p);
}

int sqk_segmenter_pa(sqk_ctx *c, const int16_t *signals, const int64_t *offsets, int64_t n_reads, int64_t max_read_len,
                     const double *pa_offset, const double *pa_scale, const sqk_seg_params *p, int mem, int32_t *segs,
                     int32_t *n_segs)
{
    if (!pa_offset || !pa_scale) return fail(SQK_ERR_ARG, "pa_offset/pa_scale is NULL");
    return segmenter_impl(c, signals, offsets, n_reads, max_read_len, pa_offset, pa_scale, p, mem, segs, n_segs);
}

int sqk_motifseq_trace(sqk_ctx *c, const int16_t *signal, int64_t n_samples, const double *model, int32_t n_model,
                       const sqk_motif_params *p, double *last_row, double *norm_sig, int64_t cap, int64_t *n_out,
                       sqk_hit *hit)
{
    if (!c || !signal || !model || !n_out) return fail(SQK_ERR_ARG, "NULL argument");
    if (n_samples < 1 || n_samples > 0x7fffffffLL) return fail(SQK_ERR_ARG, "n_samples out of range");
    if (n_model < 1 || n_model > 1024) return fail(SQK_ERR_UNSUPPORTED, "trace supports motifs of 1..1024 points");
    TRY(check_motif_params(p));
    Guard g(c->device);
    Slot &s = c->slot[0];
    cudaStream_t st = s.stream;
    CU(cudaStreamSynchronize(st));
    const int64_t offs[2] = {0, n_samples};
    TRY(ensure(s.signals, (size_t)n_samples * 2 + 16));
    TRY(ensure(s.offsets, 2 * sizeof(int64_t)));
    TRY(ensure(s.hits, sizeof(sqk_hit)));
    TRY(ensure(s.nkept, sizeof(int32_t)));
    TRY(host_begin(c));
    TRY(upload_models(c, model, (size_t)n_model, st));
    CU(cudaMemcpyAsync(s.signals.p, signal, (size_t)n_samples * 2, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(s.offsets.p, offs, sizeof(offs), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));   // offs lives on this stack frame
    View v{(const int16_t *)s.signals.p, 0, n_samples, (const int64_t *)s.offsets.p, 0, 1, n_samples};
    const int32_t mo[2] = {0, n_model};
    TRY(enqueue_motifseq(c, s, st, v, (const double *)c->model.p, model, mo, 1, p, (sqk_hit *)s.hits.p, (int32_t *)s.nkept.p));
    ReadStats rs;
    sqk_hit h;
    CU(cudaMemcpyAsync(&rs, s.stats.p, sizeof(rs), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(&h, s.hits.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (hit) *hit = h;
    *n_out = rs.n_kept;
    if ((last_row || norm_sig) && rs.n_kept > 0 && !(rs.flags & SQK_FLAG_DEGENERATE)) {
        if (cap < rs.n_kept) return fail(SQK_ERR_OVERFLOW, "cap %lld < %d kept samples", (long long)cap, rs.n_kept);
        DevBuf ybuf, rowbuf, nbuf;
        int rc = ensure(ybuf, (size_t)n_samples * sizeof(double));
        if (rc == SQK_OK) rc = ensure(rowbuf, (size_t)n_samples * sizeof(double));
        if (rc == SQK_OK) rc = ensure(nbuf, sizeof(int));
        if (rc != SQK_OK) { release(ybuf); release(rowbuf); release(nbuf); return rc; }
        sqk_trace_normalise_kernel<<<1, 256, 0, st>>>((const int16_t *)s.signals.p, n_samples, p->lo, p->hi, rs.center,
                                                     rs.scale, (double *)ybuf.p, (int *)nbuf.p);
        if (last_row) {
            const int threads = ((n_model + 31) / 32) * 32;
            sqk_trace_lastrow_kernel<<<1, threads, 2 * n_model * sizeof(double), st>>>(
                (const double *)ybuf.p, rs.n_kept, (const double *)c->model.p, n_model, (double *)rowbuf.p);
            cudaMemcpyAsync(last_row, rowbuf.p, (size_t)rs.n_kept * sizeof(double), cudaMemcpyDeviceToHost, st);
        }
        if (norm_sig) cudaMemcpyAsync(norm_sig, ybuf.p, (size_t)rs.n_kept * sizeof(double), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        release(ybuf); release(rowbuf); release(nbuf);
        if (e != cudaSuccess) return fail(SQK_ERR_CUDA, "trace kernels: %s", cudaGetErrorString(e));
    }
    return SQK_OK;
}

}  // extern "C"
