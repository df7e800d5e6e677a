// sqk_ubench.cu -- register-only micro-benchmarks that define the ALU roofline of the DTW kernel
// ("alu_fraction = cells/s / measured cells/s of a register-only micro-benchmark of
// the same instruction mix").  Measurement tool, not part of libsqk.so.
//
//   dtw_step<T,K,L>  the exact inner step of sqk_dtw_kernel (same code, included below) run on every
//                    lane with no global memory traffic and no refill: upper bound for the kernel.
//   pipe tests       issue rates of the individual SASS ops the step is made of (DADD, DSETP, FSEL,
//                    SEL, SHFL) so the binding pipe can be named.
// Prints one JSON object per line.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sqk_dtw_experiments.cuh"
#include "sqk_dtw_lb.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <typename T, int K, int L>
__global__ void __launch_bounds__(SQK_DTW_THREADS) ub_dtw(int steps, const double *model, T *sink)
{
    constexpr int G = 32 / L, RC = 16 * L;
    __shared__ T ring_all[SQK_DTW_WARPS * G * RC];
    const int lane = threadIdx.x & 31, l = lane % L, g = lane / L;
    T *ring = ring_all + ((threadIdx.x >> 5) * G + g) * RC;
    for (int q = l; q < RC; q += L) ring[q] = (T)(((q * 2654435761u) >> 20) & 1023) * (T)(1.0 / 256) - (T)2;
    __syncwarp();
    T x[K], c[K], c2[K];
    int s[K], s2[K];
#pragma unroll
    for (int k = 0; k < K; k++) { x[k] = (T)model[(l * K + k) % 80]; c[k] = DtwNum<T>::inf(); s[k] = 0; }
    T bot_c = DtwNum<T>::inf(), prev_up_c = (l == 0) ? (T)0 : DtwNum<T>::inf(), best = DtwNum<T>::inf();
    int bot_s = 0, prev_up_s = 0, best_j = -1, best_s = -1;
    const int n = (l == L - 1) ? steps : 0;
    for (int t = 0; t < steps; t += 2) {
        dtw_step<T, K, L, false, false>(c, s, c2, s2, x, ring, l, false, t, n, bot_c, bot_s, prev_up_c, prev_up_s, best, best_j, best_s, 0, false);
        dtw_step<T, K, L, false, false>(c2, s2, c, s, x, ring, l, false, t + 1, n, bot_c, bot_s, prev_up_c, prev_up_s, best, best_j, best_s, 0, false);
    }
    T acc = best + (T)best_j + (T)best_s;
#pragma unroll
    for (int k = 0; k < K; k++) acc += c[k] + (T)s[k];
    if (acc == (T)123456.789) sink[0] = acc;
}

template <typename T, int K, int L>
__global__ void __launch_bounds__(SQK_DTW_THREADS) ub_dtw_cost(int steps, const double *model, T *sink)
{
    constexpr int G = 32 / L, RC = 16 * L;
    __shared__ T ring_all[SQK_DTW_WARPS * G * RC];
    const int lane = threadIdx.x & 31, l = lane % L, g = lane / L;
    T *ring = ring_all + ((threadIdx.x >> 5) * G + g) * RC;
    for (int q = l; q < RC; q += L) ring[q] = (T)(((q * 2654435761u) >> 20) & 1023) * (T)(1.0 / 256) - (T)2;
    __syncwarp();
    T x[K], c[K], c2[K];
#pragma unroll
    for (int k = 0; k < K; k++) { x[k] = (T)model[(l * K + k) % 80]; c[k] = DtwNum<T>::inf(); }
    T bot_c = DtwNum<T>::inf(), prev_up_c = (l == 0) ? (T)0 : DtwNum<T>::inf(), best = DtwNum<T>::inf();
    int best_j = -1;
    const int n = (l == L - 1) ? steps : 0;
    for (int t = 0; t < steps; t += 2) {
        dtw_step_cost<T, K, L, false>(c, c2, x, ring, l, false, t, n, bot_c, prev_up_c, best, best_j);
        dtw_step_cost<T, K, L, false>(c2, c, x, ring, l, false, t + 1, n, bot_c, prev_up_c, best, best_j);
    }
    T acc = best + (T)best_j;
#pragma unroll
    for (int k = 0; k < K; k++) acc += c[k];
    if (acc == (T)123456.789) sink[0] = acc;
}

template <typename T, int K, int L>
__global__ void __launch_bounds__(SQK_DTW_THREADS) ub_dtw2(int steps, const double *model, T *sink)
{
    constexpr int G = 32 / L, RC = 16 * L;
    __shared__ __align__(16) T ring_all[SQK_DTW_WARPS * G * RC];
    const int lane = threadIdx.x & 31, l = lane % L, g = lane / L;
    T *ring = ring_all + ((threadIdx.x >> 5) * G + g) * RC;
    for (int q = l; q < RC; q += L) ring[q] = (T)(((q * 2654435761u) >> 20) & 1023) * (T)(1.0 / 256) - (T)2;
    __syncwarp();
    T x[K], c[K];
    int s[K];
#pragma unroll
    for (int k = 0; k < K; k++) { x[k] = (T)model[(l * K + k) % 80]; c[k] = DtwNum<T>::inf(); s[k] = 0; }
    T botA = DtwNum<T>::inf(), botB = DtwNum<T>::inf(), prev_c = (l == 0) ? (T)0 : DtwNum<T>::inf(), best = DtwNum<T>::inf();
    int botA_s = 0, botB_s = 0, prev_s = 0, best_j = -1, best_s = -1;
    const int n = (l == L - 1) ? steps : 0;
    for (int tp = 0; tp < steps / 2; tp++)
        dtw_step2<T, K, L, false>(c, s, x, ring, l, false, tp, n, botA, botA_s, botB, botB_s, prev_c, prev_s, best, best_j, best_s);
    T acc = best + (T)best_j + (T)best_s;
#pragma unroll
    for (int k = 0; k < K; k++) acc += c[k] + (T)s[k];
    if (acc == (T)123456.789) sink[0] = acc;
}

// the inner step of the float32 lower-bound kernel (pass 1 of the two-pass plan), register/shared-memory only
template <int K, int L>
__global__ void __launch_bounds__(SQK_LB_THREADS, SQK_LB_MINB(K)) ub_lb(int steps, const double *model, float *sink)
{
    constexpr int G = 32 / L, RC = 16 * L;
    __shared__ float ring_raw[SQK_LB_WARPS * G * RC + RC];
    __shared__ LbClusters cl_all[SQK_LB_WARPS * G];
    __shared__ int32_t ck_all[SQK_LB_WARPS * G * SQK_LB_CKPT];
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(ring_raw);
    float *ring_all = ring_raw + (((s0 + RC * 4u - 1u) & ~(RC * 4u - 1u)) - s0) / 4u;
    const int lane = threadIdx.x & 31, l = lane % L, g = lane / L;
    const int gid = (threadIdx.x >> 5) * G + g;
    float *ring = ring_all + gid * RC;
    LbClusters *cl = cl_all + gid;
    int32_t *ck = ck_all + gid * SQK_LB_CKPT;
    for (int q = l; q < SQK_LB_CKPT; q += L) ck[q] = q * 60;
    for (int q = l; q < RC; q += L) ring[q] = (float)(((q * 2654435761u) >> 20) & 1023) * (1.0f / 256) - 2.0f;
    if (l == L - 1) lbc_reset(*cl);
    __syncwarp();
    float x[K], c[K], c2[K];
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < K; k++) { x[k] = (float)model[(l * K + k) % 80]; c[k] = inf; }
    float bot = inf, prev_up = (l == 0) ? 0.0f : inf, tf = 0.0f;
    LbWatch wt; wt.runmin = inf; wt.thr = SQK_LB_THR_INIT; wt.thr_u = -inf; wt.n = steps; wt.N = 80;
    wt.w = sqk_lb_width(2.0, 8.0);
    sqk_lb_slack(80, wt.w, &wt.aeps, &wt.bslack);
    unsigned raddr = (unsigned)__cvta_generic_to_shared(ring) + 4u * (unsigned)((g * L - l) & (RC - 1));
    for (int t = 0; t < steps; t += 2) {
        if ((t % (7 * L)) == 0) { if (l == L - 1) wt.thr_u = sqk_lb_thr_u(wt.thr, sqk_mul_ru((float)(t + 7 * L + 80), wt.w)); tf = sqk_lb_virtual((float)t, wt.w); }
        lb_step<K, L, false>(c, c2, x, raddr, l, false, t, tf, bot, prev_up, wt, cl, ck, t / 64, 0, 192);
        lb_step<K, L, false>(c2, c, x, raddr, l, false, t + 1, tf, bot, prev_up, wt, cl, ck, t / 64, 0, 192);
    }
    float acc = wt.runmin + wt.thr + (float)cl->n;
#pragma unroll
    for (int k = 0; k < K; k++) acc += c[k];
    if (acc == 123456.789f) sink[0] = acc;
}

// two columns per step (lb_step2)
template <int K, int L>
__global__ void __launch_bounds__(SQK_LB_THREADS, SQK_LB_MINB(K)) ub_lb2(int steps, const double *model, float *sink)
{
    constexpr int G = 32 / L, RC = 16 * L;
    __shared__ float ring_raw[SQK_LB_WARPS * G * RC + RC];
    __shared__ LbClusters cl_all[SQK_LB_WARPS * G];
    __shared__ int32_t ck_all[SQK_LB_WARPS * G * SQK_LB_CKPT];
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(ring_raw);
    float *ring_all = ring_raw + (((s0 + RC * 4u - 1u) & ~(RC * 4u - 1u)) - s0) / 4u;
    const int lane = threadIdx.x & 31, l = lane % L, g = lane / L;
    const int gid = (threadIdx.x >> 5) * G + g;
    float *ring = ring_all + gid * RC;
    LbClusters *cl = cl_all + gid;
    int32_t *ck = ck_all + gid * SQK_LB_CKPT;
    for (int q = l; q < SQK_LB_CKPT; q += L) ck[q] = q * 60;
    for (int q = l; q < RC; q += L) ring[q] = (float)(((q * 2654435761u) >> 20) & 1023) * (1.0f / 256) - 2.0f;
    if (l == L - 1) lbc_reset(*cl);
    __syncwarp();
    float x[K], c[K], c2[K];
    const float inf = __int_as_float(0x7f800000);
#pragma unroll
    for (int k = 0; k < K; k++) { x[k] = (float)model[(l * K + k) % 80]; c[k] = inf; }
    float bot_a = inf, bot_b = inf, prev_up_b = (l == 0) ? 0.0f : inf, tf = 0.0f;
    LbWatch wt; wt.runmin = inf; wt.thr = SQK_LB_THR_INIT; wt.thr_u = -inf; wt.n = steps; wt.N = 80;
    wt.w = sqk_lb_width(2.0, 8.0);
    sqk_lb_slack(80, wt.w, &wt.aeps, &wt.bslack);
    unsigned raddr = (unsigned)__cvta_generic_to_shared(ring) + 4u * (unsigned)((g * L - 2 * l) & (RC - 1));
    for (int t = 0; t < steps; t += 4) {
        if ((t % (6 * L)) == 0) { if (l == L - 1) wt.thr_u = sqk_lb_thr_u(wt.thr, sqk_mul_ru((float)(t + 6 * L + 80), wt.w)); tf = sqk_lb_virtual((float)t, wt.w); }
        lb_step2<K, L, false>(c, c2, x, raddr, l, false, t, tf, bot_a, bot_b, prev_up_b, wt, cl, ck, t / 64, 0, 192);
        lb_step2<K, L, false>(c2, c, x, raddr, l, false, t + 2, tf, bot_a, bot_b, prev_up_b, wt, cl, ck, t / 64, 0, 192);
    }
    float acc = wt.runmin + wt.thr + (float)cl->n;
#pragma unroll
    for (int k = 0; k < K; k++) acc += c[k];
    if (acc == 123456.789f) sink[0] = acc;
}

template <int K, int L>
static void run_lb2(int sms, const double *d_model, void *d_sink)
{
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ub_lb2<K, L>, SQK_LB_THREADS, 0));
    const int grid = sms * occ, steps = 8192;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    ub_lb2<K, L><<<grid, SQK_LB_THREADS>>>(256, d_model, (float *)d_sink);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a));
        ub_lb2<K, L><<<grid, SQK_LB_THREADS>>>(steps, d_model, (float *)d_sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)grid * SQK_LB_THREADS * K * steps;
    printf("{\"bench\": \"lb_step2\", \"precision\": \"fp32_rd\", \"K\": %d, \"L\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, "
           "\"cells_per_s\": %.4e}\n", K, L, occ, best_ms, cells / (best_ms * 1e-3));
    fflush(stdout);
}

template <int K, int L>
static void run_lb(int sms, const double *d_model, void *d_sink)
{
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ub_lb<K, L>, SQK_LB_THREADS, 0));
    const int grid = sms * occ, steps = 8192;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    ub_lb<K, L><<<grid, SQK_LB_THREADS>>>(256, d_model, (float *)d_sink);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a));
        ub_lb<K, L><<<grid, SQK_LB_THREADS>>>(steps, d_model, (float *)d_sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)grid * SQK_LB_THREADS * K * steps;
    printf("{\"bench\": \"lb_step\", \"precision\": \"fp32_rd\", \"K\": %d, \"L\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, "
           "\"cells_per_s\": %.4e}\n", K, L, occ, best_ms, cells / (best_ms * 1e-3));
    fflush(stdout);
}

template <typename T, int K, int L>
static void run_dtw2(const char *prec, int sms, const double *d_model, void *d_sink)
{
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ub_dtw2<T, K, L>, SQK_DTW_THREADS, 0));
    const int grid = sms * occ, steps = 8192;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    ub_dtw2<T, K, L><<<grid, SQK_DTW_THREADS>>>(256, d_model, (T *)d_sink);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a));
        ub_dtw2<T, K, L><<<grid, SQK_DTW_THREADS>>>(steps, d_model, (T *)d_sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)grid * SQK_DTW_THREADS * K * steps;
    printf("{\"bench\": \"dtw_step2\", \"precision\": \"%s\", \"K\": %d, \"L\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, "
           "\"cells_per_s\": %.4e}\n", prec, K, L, occ, best_ms, cells / (best_ms * 1e-3));
    fflush(stdout);
}

template <typename T, int K, int L>
static void run_dtw_cost(const char *prec, int sms, const double *d_model, void *d_sink)
{
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ub_dtw_cost<T, K, L>, SQK_DTW_THREADS, 0));
    const int grid = sms * occ, steps = 8192;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    ub_dtw_cost<T, K, L><<<grid, SQK_DTW_THREADS>>>(256, d_model, (T *)d_sink);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a));
        ub_dtw_cost<T, K, L><<<grid, SQK_DTW_THREADS>>>(steps, d_model, (T *)d_sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)grid * SQK_DTW_THREADS * K * steps;
    printf("{\"bench\": \"dtw_step_cost_only\", \"precision\": \"%s\", \"K\": %d, \"L\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, "
           "\"cells_per_s\": %.4e}\n", prec, K, L, occ, best_ms, cells / (best_ms * 1e-3));
    fflush(stdout);
}

template <typename T, int K, int L>
static void run_dtw(const char *prec, int sms, const double *d_model, void *d_sink)
{
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ub_dtw<T, K, L>, SQK_DTW_THREADS, 0));
    const int grid = sms * occ, steps = 8192;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    ub_dtw<T, K, L><<<grid, SQK_DTW_THREADS>>>(256, d_model, (T *)d_sink);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a));
        ub_dtw<T, K, L><<<grid, SQK_DTW_THREADS>>>(steps, d_model, (T *)d_sink);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    const double cells = (double)grid * SQK_DTW_THREADS * K * steps;
    printf("{\"bench\": \"dtw_step\", \"precision\": \"%s\", \"K\": %d, \"L\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, "
           "\"cells_per_s\": %.4e}\n", prec, K, L, occ, best_ms, cells / (best_ms * 1e-3));
    fflush(stdout);
}

// ---- individual pipes: ILP independent chains per thread, enough warps to saturate ----------------
enum { OP_DADD, OP_DSETP_FSEL, OP_FSEL, OP_SEL, OP_SHFL, OP_FADD, OP_IMAD, OP_FMNMX, OP_FADD_RM, OP_FADD_RZ_ABS, OP_FMNMX3, OP_LBCELL };

template <int OP>
__global__ void __launch_bounds__(256) ub_pipe(int iters, double *sink, int seed)
{
    constexpr int C = 8;
    double d[C]; float f[C]; int v[C];
#pragma unroll
    for (int i = 0; i < C; i++) { d[i] = threadIdx.x * 0.001 + i + seed; f[i] = (float)d[i]; v[i] = threadIdx.x + i * seed; }
    const double inc = 1.0 + seed * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < C; i++) {
            if (OP == OP_DADD) d[i] = __dadd_rn(d[i], inc);
            if (OP == OP_DSETP_FSEL) d[i] = (d[i] < d[(i + 1) % C]) ? d[i] : d[(i + 3) % C];
            if (OP == OP_FSEL) f[i] = (v[i] > it) ? f[i] : f[(i + 3) % C];
            if (OP == OP_SEL) v[i] = (v[(i + 1) % C] > it) ? v[i] : v[(i + 3) % C];
            if (OP == OP_SHFL) v[i] = __shfl_up_sync(0xffffffffu, v[i], 1, 8);
            if (OP == OP_FADD) f[i] = __fadd_rn(f[i], (float)inc);
            if (OP == OP_IMAD) v[i] = v[i] * seed + it;
            if (OP == OP_FMNMX) f[i] = fminf(f[i], f[(i + 3) % C] + 0.0f);
            if (OP == OP_FADD_RM) f[i] = __fadd_rd(f[i], (float)inc);
            if (OP == OP_FADD_RZ_ABS) f[i] = __fadd_rz(fabsf(f[i]), -(float)inc);
            if (OP == OP_FMNMX3) f[i] = fminf(fminf(f[i], f[(i + 3) % C]), f[(i + 5) % C]);
            if (OP == OP_LBCELL) f[i] = __fadd_rd(__fadd_rd(fabsf(__fadd_rz(f[(i + 1) % C], -(float)inc)), -1e-7f), fminf(fminf(f[i], f[(i + 3) % C]), f[(i + 5) % C]));
        }
    }
    double acc = 0;
#pragma unroll
    for (int i = 0; i < C; i++) acc += d[i] + f[i] + v[i];
    if (acc == 123456.789) sink[0] = acc;
}

template <int OP>
static void run_pipe(const char *name, int sms, int clock_khz, void *d_sink)
{
    const int grid = sms * 8, iters = 4096;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    ub_pipe<OP><<<grid, 256>>>(64, (double *)d_sink, 1);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(a));
        ub_pipe<OP><<<grid, 256>>>(iters, (double *)d_sink, 1);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best_ms) best_ms = ms;
    }
    const double ops = (double)grid * 256 * 8 * iters;   // thread-level source ops
    const double per_sm_clk = ops / (best_ms * 1e-3) / sms / (clock_khz * 1e3);
    printf("{\"bench\": \"pipe\", \"op\": \"%s\", \"ms\": %.4f, \"thread_ops_per_s\": %.4e, \"per_sm_per_clk_at_max_clock\": %.2f}\n",
           name, best_ms, ops / (best_ms * 1e-3), per_sm_clk);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const bool full = argc > 1;
    int dev = 0, sms = 0, clk = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev));
    std::vector<double> model(80);
    for (int i = 0; i < 80; i++) model[i] = ((i * 37) % 17) * 0.25 - 2.0;
    double *d_model; void *d_sink;
    CK(cudaMalloc(&d_model, 80 * sizeof(double)));
    CK(cudaMalloc(&d_sink, 64));
    CK(cudaMemcpy(d_model, model.data(), 80 * sizeof(double), cudaMemcpyHostToDevice));
    printf("{\"bench\": \"device\", \"sms\": %d, \"clock_khz\": %d}\n", sms, clk);
    run_lb2<10, 8>(sms, d_model, d_sink);
    run_lb2<20, 4>(sms, d_model, d_sink);
    run_lb2<5, 16>(sms, d_model, d_sink);
    run_lb<10, 8>(sms, d_model, d_sink);
    run_lb<20, 4>(sms, d_model, d_sink);
    run_lb<5, 16>(sms, d_model, d_sink);
    run_dtw<double, 20, 4>("fp64", sms, d_model, d_sink);
    run_dtw<double, 10, 8>("fp64", sms, d_model, d_sink);
    run_dtw<double, 5, 16>("fp64", sms, d_model, d_sink);
    run_dtw<float, 20, 4>("fp32", sms, d_model, d_sink);
    run_dtw<float, 10, 8>("fp32", sms, d_model, d_sink);
    run_dtw2<double, 10, 8>("fp64", sms, d_model, d_sink);
    run_dtw2<double, 20, 4>("fp64", sms, d_model, d_sink);
    run_dtw2<double, 5, 16>("fp64", sms, d_model, d_sink);
    run_dtw2<float, 10, 8>("fp32", sms, d_model, d_sink);
    run_dtw_cost<double, 10, 8>("fp64", sms, d_model, d_sink);
    run_dtw_cost<double, 20, 4>("fp64", sms, d_model, d_sink);
    if (full) {
        run_dtw_cost<double, 5, 16>("fp64", sms, d_model, d_sink);
        run_dtw_cost<float, 10, 8>("fp32", sms, d_model, d_sink);
        run_dtw<double, 3, 32>("fp64", sms, d_model, d_sink);
        run_dtw<double, 16, 4>("fp64", sms, d_model, d_sink);
        run_dtw<double, 12, 8>("fp64", sms, d_model, d_sink);
        run_pipe<OP_DADD>("DADD", sms, clk, d_sink);
        run_pipe<OP_DSETP_FSEL>("DSETP+2FSEL", sms, clk, d_sink);
        run_pipe<OP_FSEL>("FSEL", sms, clk, d_sink);
        run_pipe<OP_SEL>("ISETP+SEL", sms, clk, d_sink);
        run_pipe<OP_SHFL>("SHFL", sms, clk, d_sink);
        run_pipe<OP_FADD>("FADD", sms, clk, d_sink);
        run_pipe<OP_IMAD>("IMAD", sms, clk, d_sink);
        run_pipe<OP_FMNMX>("FADD+FMNMX", sms, clk, d_sink);
        run_pipe<OP_FADD_RM>("FADD.RM", sms, clk, d_sink);
        run_pipe<OP_FADD_RZ_ABS>("FADD.RZ|abs|", sms, clk, d_sink);
        run_pipe<OP_FMNMX3>("FMNMX3", sms, clk, d_sink);
        run_pipe<OP_LBCELL>("lb cell (FADD.RZ+FADD.RM+FMNMX3+FADD.RM), 8 independent chains", sms, clk, d_sink);
    }
    return 0;
}
