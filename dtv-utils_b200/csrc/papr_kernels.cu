// papr_kernels.cu — hand-written sm_100a kernels of the PAPR/CCDF hot path.
//
// What the reference does per sample (drmpeg/dtv-utils papr.c):
//   pass 1  papr.c:100-129   v = I*I + Q*Q (three separately rounded float32 ops), sum += v (double),
//                            strict-> first-occurrence maxima of v, +I, -I, +Q, -Q
//   pass 2  papr.c:143-153   for every level j: if (v > level[j]) level_count[j]++      (O(N*L))
//           papr.c:175-185   same with 0.1 dB levels (-g)
//
// How it is done here (HBM-bound byte/float work, no tensor cores):
//   * one persistent grid (1 CTA x 1024 threads per SM); a warp owns 2 KiB-contiguous "batches"
//     (32 lanes x 4 x 16-byte streaming loads in flight per lane), batches ascend per warp so a
//     strict compare keeps the first occurrence;
//   * extremes: per-lane FMNMX into batch maxima, one REDUX.MAX per tracker per batch, and only on the
//     (warp-uniform, rare) improvement a REDUX.MIN locates the exact sample - no per-lane index state;
//   * sum: per-lane double accumulation, fixed-order shuffle tree, per-warp partials reduced in a
//     fixed order by a 1-CTA follow-up kernel (deterministic run to run);
//   * CCDF: the O(N*L) compare loop is replaced by a histogram of "cells" (float32 bit pattern >> sh,
//     i.e. 2048 log-spaced cells per octave) in shared memory; only samples whose cell contains a
//     threshold (or, in the fused mode, may contain it) are additionally counted exactly, one counter
//     per float32 value, in a small L2-resident fine table.  level_count[j] = samples in cells above
//     the threshold's cell + the fine counters above the threshold inside its cell.  Exact.
#include "papr_scan_common.cuh"

template <bool STATS, bool HIST>
__global__ void __launch_bounds__(PAPR_THREADS, PAPR_CTAS_PER_SM) papr_scan_kernel(const PaprScanArgs a)
{
    ScanState<STATS, HIST> st;
    unsigned *s_hist = reinterpret_cast<unsigned *>(scan_smem);
    const int lane = threadIdx.x & 31;
    const unsigned gw = blockIdx.x * PAPR_WARPS + (threadIdx.x >> 5);
    const unsigned GW = gridDim.x * PAPR_WARPS;
    bool do_hist = false;

    if (HIST) {
        const PaprPlan pl = *a.plan;
        do_hist = (pl.status & PLAN_HIST) != 0;
        if (!STATS && !do_hist) return;
        unsigned *s_fb = s_hist + (PAPR_NCELLS_MAX + 2);
        st.g_fine = a.g_fine;
        // opaque to the optimiser on purpose: otherwise the window base is rematerialised per sample
        asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}"
                     : "=r"(st.smem_slot0) : "l"(s_hist));
        st.sh = pl.sh;
        // without a valid plan (fused mode, nothing to predict from) every sample lands in slot 0
        st.neg_base1 = 1 - (do_hist ? pl.cell_base : 0x7fffffff);
        st.ncells = do_hist ? pl.ncells : 0;
        st.fmask = (1u << pl.sh) - 1u;
        for (int i = threadIdx.x; i < st.ncells + 2; i += PAPR_THREADS) {
            s_hist[i] = 0;
            s_fb[i] = (i >= 1 && i <= st.ncells) ? a.fine_base[i - 1] : 0u;
        }
        __syncthreads();
    }
    if (STATS) {
        st.dsum = 0.0;
        st.upd = 0;
#pragma unroll
        for (int t = 0; t < PAPR_NTRACK; ++t) {
            st.run_val[t] = a.wp[blockIdx.x].val[t]; // carried across the chunk launches of a streamed shard
            st.run_pos[t] = 0;
        }
    }

    const float4 *p = reinterpret_cast<const float4 *>(a.iq);
    const unsigned nbatch_full = (unsigned)(a.nsamples / PAPR_BATCH_SAMPLES);
    for (unsigned b = gw; b < nbatch_full; b += GW) {
        const float4 *q = p + (size_t)b * PAPR_BATCH_VEC + lane;
        float4 r[PAPR_U];
#pragma unroll
        for (int u = 0; u < PAPR_U; ++u) r[u] = ldg_stream(q + 32 * u);
        process_batch(st, r, b * PAPR_BATCH_SAMPLES, lane);
    }
    // ragged last batch (zero padded: zeros add +0.0 to the sum, beat no maximum, exceed no threshold)
    const u64 done = (u64)nbatch_full * PAPR_BATCH_SAMPLES;
    if (done < a.nsamples && gw == nbatch_full % GW) {
        float4 r[PAPR_U];
#pragma unroll
        for (int u = 0; u < PAPR_U; ++u) {
            u64 s0 = done + 2u * (unsigned)(u * 32 + lane);
            r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s0 + 1 < a.nsamples) {
                r[u] = ldg_stream(p + (s0 >> 1));
            } else if (s0 < a.nsamples) {
                float2 t = *reinterpret_cast<const float2 *>(a.iq + 2 * s0);
                r[u].x = t.x;
                r[u].y = t.y;
            }
        }
        process_batch(st, r, (unsigned)done, lane);
    }

    if (STATS) {
        // CTA-level fold of the warps' states (fixed order), then into the CTA's persistent partial
        __shared__ double s_wsum[PAPR_WARPS];
        __shared__ int s_wval[PAPR_NTRACK][PAPR_WARPS];
        __shared__ unsigned s_wpos[PAPR_NTRACK][PAPR_WARPS];
        const int warp = threadIdx.x >> 5;
        double wsum = warp_sum_fixed(st.dsum);
        if (lane == 0) {
            s_wsum[warp] = wsum;
#pragma unroll
            for (int t = 0; t < PAPR_NTRACK; ++t) {
                const bool u = (st.upd >> t) & 1u;
                s_wval[t][warp] = u ? st.run_val[t] : 0; // 0 = nothing new from this warp
                s_wpos[t][warp] = u ? st.run_pos[t] : 0xffffffffu;
            }
        }
        __syncthreads();
        if (warp == 0) {
            double x = lane < PAPR_WARPS ? s_wsum[lane] : 0.0;
            x = warp_sum_fixed(x);
            PaprCtaPartial *w = a.wp + blockIdx.x;
            if (lane == 0) w->sum += x;
#pragma unroll
            for (int t = 0; t < PAPR_NTRACK; ++t) {
                int v = lane < PAPR_WARPS ? s_wval[t][lane] : 0;
                unsigned pos = lane < PAPR_WARPS ? s_wpos[t][lane] : 0xffffffffu;
                const int vmax = __reduce_max_sync(FULL, v);
                const unsigned pmin = __reduce_min_sync(FULL, v == vmax ? pos : 0xffffffffu);
                // strictly greater than what earlier launches (= lower indices) left: first occurrence kept
                if (lane == 0 && vmax > w->val[t]) {
                    w->val[t] = vmax;
                    w->idx[t] = a.first_index + pmin;
                }
            }
        }
    }
    if (HIST && do_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < st.ncells; i += PAPR_THREADS) {
            unsigned c = s_hist[i + 1];
            if (c) atomicAdd(&a.g_hist[i], (u64)c);
        }
        if (threadIdx.x == 0 && s_hist[st.ncells + 1]) atomicAdd(a.g_over, (u64)s_hist[st.ncells + 1]);
    }
}

int papr_scan_smem_bytes(bool hist) { return hist ? SCAN_SMEM_BYTES : 0; }

int papr_scan_configure(void)
{
    int smem = papr_scan_smem_bytes(true);
    cudaError_t e1 = cudaFuncSetAttribute(papr_scan_kernel<true, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaError_t e2 = cudaFuncSetAttribute(papr_scan_kernel<false, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    return (e1 == cudaSuccess && e2 == cudaSuccess) ? 0 : -1;
}

void papr_launch_scan(bool stats, bool hist, int grid, const PaprScanArgs &a, cudaStream_t s)
{
    if (stats && hist)
        papr_scan_kernel<true, true><<<grid, PAPR_THREADS, papr_scan_smem_bytes(true), s>>>(a);
    else if (stats)
        papr_scan_kernel<true, false><<<grid, PAPR_THREADS, 0, s>>>(a);
    else
        papr_scan_kernel<false, true><<<grid, PAPR_THREADS, papr_scan_smem_bytes(true), s>>>(a);
}

// ------------------------------------------------------------------------------------------------
// pass-1 follow-ups
// ------------------------------------------------------------------------------------------------
#include "papr_finalize.cuh"

// one CTA.  with_levels: single-shard analysis, the local state IS the merged state.
__global__ void __launch_bounds__(FIN_T) papr_stats_finalize_kernel(const PaprCtaPartial *wp, int nctas, u64 n,
                                                                   PaprDevStats *out, bool with_levels,
                                                                   PaprTables tb, int graph, PaprDevStats *merged,
                                                                   PaprDevLevels *lv, u64 *status_word,
                                                                   const PaprChainList *chain, int *chain_report)
{
    reduce_partials(wp, nctas, n, out);
    if (chain && threadIdx.x == 0) { // the sequential sum chained on the device replaces the approximate one
        const int st = chain->status;
        if (st == XT_OK) out->sum = chain->exact;
        chain_report[0] = st;
        chain_report[1] = chain->why;
    }
    if (!with_levels) return;
    __syncthreads(); // thread 0's global writes are read back by thread 0 only
    merge_and_levels(out, 1, tb, graph, merged, lv, status_word);
}

void papr_launch_stats_finalize(const PaprCtaPartial *wp, int nctas, u64 n, PaprDevStats *out, cudaStream_t s)
{
    PaprTables none = {nullptr, nullptr, 0};
    papr_stats_finalize_kernel<<<1, FIN_T, 0, s>>>(wp, nctas, n, out, false, none, 0, nullptr, nullptr, nullptr, nullptr, nullptr);
}

void papr_launch_finalize_levels(const PaprCtaPartial *wp, int nctas, u64 n, PaprTables t, int graph,
                                 PaprDevStats *local, PaprDevStats *merged, PaprDevLevels *lv,
                                 u64 *status_word, const PaprChainList *chain, int *chain_report, cudaStream_t s)
{
    papr_stats_finalize_kernel<<<1, FIN_T, 0, s>>>(wp, nctas, n, local, true, t, graph, merged, lv, status_word, chain,
                                                   chain_report);
}

__global__ void __launch_bounds__(FIN_T) papr_levels_kernel(const PaprDevStats *parts, int nparts, PaprTables tb,
                                                           int graph, PaprDevStats *merged, PaprDevLevels *lv,
                                                           u64 *status_word, const PaprChainList *chain, int *chain_report)
{
    merge_and_levels(parts, nparts, tb, graph, merged, lv, status_word, sizeof(PaprDevStats), chain, chain_report);
}

void papr_launch_levels(const PaprDevStats *parts, int nparts, PaprTables t, int graph,
                        PaprDevStats *merged, PaprDevLevels *lv, u64 *status_word, cudaStream_t s,
                        const PaprChainList *chain, int *chain_report)
{
    papr_levels_kernel<<<1, FIN_T, 0, s>>>(parts, nparts, t, graph, merged, lv, status_word, chain, chain_report);
}

// ------------------------------------------------------------------------------------------------
// fused mode: strided subsample -> predicted mean and its standard error
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned hash32(unsigned x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// Every `stride`-th warp batch (at a hashed position inside its group, so a periodic capture cannot
// alias with the stride).  Per CTA: sum, sum of squares and count of the *batch* sums - batch means
// give an honest standard error even when neighbouring samples are correlated.
__global__ void __launch_bounds__(PAPR_THREADS, 1) papr_presample_kernel(const float *iq, u64 nsamples,
                                                                         int stride, double *cta_pre)
{
    const int lane = threadIdx.x & 31;
    const unsigned gw = blockIdx.x * PAPR_WARPS + (threadIdx.x >> 5);
    const unsigned GW = gridDim.x * PAPR_WARPS;
    const unsigned nbatch = (unsigned)(nsamples / PAPR_BATCH_SAMPLES);
    const unsigned ngroups = (nbatch + stride - 1) / stride;
    const float4 *p = reinterpret_cast<const float4 *>(iq);
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    // two groups per trip: 8 independent 16-byte loads in flight per lane (the kernel is latency-bound: a warp
    // visits only ~14 groups); the batch sums are accumulated in ascending group order, as before
    for (unsigned g0 = gw; g0 < ngroups; g0 += 2 * GW) {
        float4 r[2][PAPR_U];
        bool have[2];
#pragma unroll
        for (int k4 = 0; k4 < 2; ++k4) {
            const unsigned g = g0 + k4 * GW;
            const unsigned b = g < ngroups ? g * stride + hash32(g) % stride : nbatch;
            have[k4] = b < nbatch;
            const float4 *q = p + (size_t)(have[k4] ? b : 0) * PAPR_BATCH_VEC + lane;
#pragma unroll
            for (int u = 0; u < PAPR_U; ++u) r[k4][u] = have[k4] ? ldg_stream(q + 32 * u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k4 = 0; k4 < 2; ++k4) {
            if (!have[k4]) continue; // (warp-uniform)
            double ls = 0.0;
#pragma unroll
            for (int u = 0; u < PAPR_U; ++u) {
                ls += (double)power_of(r[k4][u].x, r[k4][u].y);
                ls += (double)power_of(r[k4][u].z, r[k4][u].w);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ls += __shfl_xor_sync(FULL, ls, o);
            s1 += ls;
            s2 += ls * ls;
            cnt += 1.0;
        }
    }
    __shared__ double s_p[3][PAPR_WARPS];
    const int warp = threadIdx.x >> 5;
    if (lane == 0) { s_p[0][warp] = s1; s_p[1][warp] = s2; s_p[2][warp] = cnt; }
    __syncthreads();
    if (warp == 0) { // fixed-order fold of the warps -> one triple per CTA
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double x = warp_sum_fixed(lane < PAPR_WARPS ? s_p[k][lane] : 0.0);
            if (lane == 0) cta_pre[3 * blockIdx.x + k] = x;
        }
    }
}

void papr_launch_presample(const float *iq, u64 nsamples, int stride, int grid, double *cta_pre,
                           cudaStream_t s)
{
    papr_presample_kernel<<<grid, PAPR_THREADS, 0, s>>>(iq, nsamples, stride, cta_pre);
}

// fixed-order fold of the CTA triples; all threads of a 1024-thread CTA return {s1, s2, cnt}
__device__ void reduce_pre(const double *cta_pre, int nctas, double out[3])
{
    __shared__ double s[3][1024];
    const int t = threadIdx.x;
    double a0 = 0, a1 = 0, a2 = 0;
    for (int i = t; i < nctas; i += 1024) {
        a0 += cta_pre[3 * i]; a1 += cta_pre[3 * i + 1]; a2 += cta_pre[3 * i + 2];
    }
    s[0][t] = a0; s[1][t] = a1; s[2][t] = a2;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (t < o) { s[0][t] += s[0][t + o]; s[1][t] += s[1][t + o]; s[2][t] += s[2][t + o]; }
        __syncthreads();
    }
    out[0] = s[0][0]; out[1] = s[1][0]; out[2] = s[2][0];
    __syncthreads();
}

__global__ void __launch_bounds__(1024) papr_presample_reduce_kernel(const double *cta_pre, int nctas, double *pre4)
{
    double r[3];
    reduce_pre(cta_pre, nctas, r);
    if (threadIdx.x == 0) { pre4[0] = r[0]; pre4[1] = r[1]; pre4[2] = r[2]; pre4[3] = 0.0; }
}

void papr_launch_presample_reduce(const double *cta_pre, int nctas, double *pre4, cudaStream_t s)
{
    papr_presample_reduce_kernel<<<1, 1024, 0, s>>>(cta_pre, nctas, pre4);
}

// ------------------------------------------------------------------------------------------------
// plans: which cells exist and which of them need exact (per-value) counting
// ------------------------------------------------------------------------------------------------
// exclusive scan of s_flag[0..PAPR_NCELLS_MAX) by one 1024-thread CTA; writes fine_base[] =
// 1 + (slot << sh), the cell's first counter in the fine table, or 0 for an unambiguous cell
__device__ int assign_slots(const unsigned char *s_flag, int ncells, int sh, int max_slots, unsigned *fine_base,
                            int *s_scan /* [1024] */)
{
    const int t = threadIdx.x;
    const int per = PAPR_NCELLS_MAX / 1024;
    int local = 0;
    for (int k = 0; k < per; ++k) {
        int c = t * per + k;
        local += (c < ncells && s_flag[c]) ? 1 : 0;
    }
    // inclusive scan over the 1024 threads: shuffle scan inside each warp, then over the 32 warp totals
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(FULL, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) s_scan[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        int w = s_scan[t];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FULL, w, o);
            if (t >= o) w += v;
        }
        s_scan[32 + t] = w; // inclusive totals of warps 0..t
    }
    __syncthreads();
    incl += (t >> 5) ? s_scan[32 + (t >> 5) - 1] : 0;
    int ord = incl - local;
    for (int k = 0; k < per; ++k) {
        int c = t * per + k;
        unsigned fb = 0;
        if (c < ncells && s_flag[c]) {
            if (ord < max_slots) fb = ((unsigned)ord << sh) + 1u;
            ord++;
        }
        fine_base[c] = fb;
    }
    int total = s_scan[63];
    __syncthreads();
    return total < max_slots ? total : max_slots;
}

// Fused mode.  pre4 = {sum, sum of squares, count} of batch sums already reduced over ranks, or
// (single shard) nctas > 0: fold this GPU's CTA triples here and skip the separate reduce launch.
__device__ void plan_pred_body(const double (&pre)[3], const PaprTables &tb, float sigmas, float bias, int fine_slots,
                               PaprPlan *plan, unsigned *fine_base, int xchg_timeout, int ncells_max)
{
    __shared__ unsigned char s_flag[PAPR_NCELLS_MAX];
    __shared__ int s_scan[1024];
    __shared__ int s_cov;
    const int t = threadIdx.x;
    const double s1 = pre[0], s2 = pre[1], cnt = pre[2];
    PaprPlan pl;
    pl.sh = PAPR_SH_MIN; pl.cell_base = 0; pl.ncells = 0; pl.n_amb = 0; pl.status = 0;
    pl.levels_covered = 0; pl.window = 0.f; pl.pad = xchg_timeout; pl.avg_pred = 0.0; pl.wcv = 0.0;
    const bool ok = cnt >= 16.0 && s1 > 0.0 && isfinite(s1) && isfinite(s2);
    if (!ok) { // nothing to predict from: the exact pass will run
        for (int c = t; c < PAPR_NCELLS_MAX; c += 1024) fine_base[c] = 0;
        if (t == 0) *plan = pl;
        return;
    }
    const double mean_b = s1 / cnt;
    double var_b = (s2 / cnt - mean_b * mean_b) * cnt / (cnt - 1.0);
    if (!(var_b > 0.0)) var_b = 0.0;
    const double se_rel = sqrt(var_b / cnt) / mean_b;
    const double w = (double)sigmas * se_rel + 0x1p-18; // floor: a few float32 ulps of the thresholds
    const double avg = mean_b / (double)PAPR_BATCH_SAMPLES * (double)bias; // bias: test hook (1.0), forces a miss
    const int sh = PAPR_SH_MIN;
    float lo0 = __double2float_rd(avg * (1.0 - w));
    if (!(lo0 > 0.f)) lo0 = 0.f;
    const int base = (int)(__float_as_uint(lo0) >> sh);
    const int ncells = min(max(ncells_max, 1024), PAPR_NCELLS_MAX); // (the TMA-fed sweep keeps a smaller histogram)
    for (int c = t; c < PAPR_NCELLS_MAX; c += 1024) s_flag[c] = 0;
    if (t == 0) s_cov = tb.nlevels_max;
    __syncthreads();
    // which levels have their whole window inside the cell range
    for (int j = t; j < tb.nlevels_max; j += 1024) {
        float hi = __double2float_ru(tb.pow10[j] * avg * (1.0 + w));
        int chi = isfinite(hi) ? (int)(__float_as_uint(hi) >> sh) - base : ncells;
        if (chi >= ncells) atomicMin(&s_cov, j);
    }
    __syncthreads();
    const int cov = s_cov;
    for (int j = t; j < cov; j += 1024) {
        float lo = __double2float_rd(tb.pow10[j] * avg * (1.0 - w));
        float hi = __double2float_ru(tb.pow10[j] * avg * (1.0 + w));
        if (!(lo > 0.f)) lo = 0.f;
        int clo = (int)(__float_as_uint(lo) >> sh) - base, chi = (int)(__float_as_uint(hi) >> sh) - base;
        if (clo < 0) clo = 0;
        for (int c = clo; c <= chi; ++c) s_flag[c] = 1;
    }
    __syncthreads();
    int n_amb = assign_slots(s_flag, ncells, sh, fine_slots, fine_base, s_scan);
    if (t == 0) {
        pl.sh = sh; pl.cell_base = base; pl.ncells = ncells; pl.n_amb = n_amb; pl.status = PLAN_HIST;
        pl.levels_covered = cov; pl.window = (float)w; pl.avg_pred = avg;
        pl.wcv = (double)sigmas * sqrt(var_b) / mean_b;
        *plan = pl;
    }
}

__global__ void __launch_bounds__(1024) papr_plan_pred_kernel(const double *pre4, const double *cta_pre, int nctas,
                                                              PaprTables tb, float sigmas, float bias, int fine_slots,
                                                              PaprPlan *plan, unsigned *fine_base, int ncells_max)
{
    double pre[3];
    if (nctas > 0) {
        reduce_pre(cta_pre, nctas, pre);
    } else {
        pre[0] = pre4[0]; pre[1] = pre4[1]; pre[2] = pre4[2];
    }
    plan_pred_body(pre, tb, sigmas, bias, fine_slots, plan, fine_base, 0, ncells_max);
}

void papr_launch_plan_pred(const double *pre4, const double *cta_pre, int nctas, PaprTables t, float sigmas, float bias,
                           int fine_slots, PaprPlan *plan, unsigned *fine_base, cudaStream_t s, int ncells_max)
{
    papr_plan_pred_kernel<<<1, 1024, 0, s>>>(pre4, cta_pre, nctas, t, sigmas, bias, fine_slots, plan, fine_base, ncells_max);
}

// Exact thresholds known (two-pass mode, or redo after a fused miss): ambiguous = the cells that hold
// a threshold.  Picks the finest cell size whose range [min level, max(level, peak)] fits.
__global__ void __launch_bounds__(1024) papr_plan_exact_kernel(const PaprDevLevels *lv, const PaprDevStats *merged,
                                                               int fine_bytes_log2, PaprPlan *plan,
                                                               unsigned *fine_base)
{
    __shared__ unsigned char s_flag[PAPR_NCELLS_MAX];
    __shared__ int s_scan[1024];
    __shared__ unsigned s_lo, s_hi;
    const int t = threadIdx.x;
    const int L = lv->L;
    PaprPlan pl;
    pl.sh = PAPR_SH_MIN; pl.cell_base = 0; pl.ncells = 0; pl.n_amb = 0; pl.status = 0;
    pl.levels_covered = L; pl.window = 0.f; pl.pad = 0; pl.avg_pred = lv->avg; pl.wcv = 0.0;
    if (L <= 0) {
        for (int c = t; c < PAPR_NCELLS_MAX; c += 1024) fine_base[c] = 0;
        if (t == 0) *plan = pl;
        return;
    }
    if (t == 0) { s_lo = 0xffffffffu; s_hi = (unsigned)merged->val[TR_PEAK]; }
    for (int c = t; c < PAPR_NCELLS_MAX; c += 1024) s_flag[c] = 0;
    __syncthreads();
    for (int j = t; j < L; j += 1024) { // levels are >= +0.0: float order == unsigned bit order
        unsigned b = __float_as_uint(lv->level[j]);
        atomicMin(&s_lo, b);
        atomicMax(&s_hi, b);
    }
    __syncthreads();
    int sh = PAPR_SH_MIN;
    while (sh < PAPR_SH_MAX && (int)((s_hi >> sh) - (s_lo >> sh)) + 1 > PAPR_NCELLS_MAX) ++sh;
    const int base = (int)(s_lo >> sh);
    const int ncells = (int)(s_hi >> sh) - base + 1;
    for (int j = t; j < L; j += 1024) s_flag[(int)(__float_as_uint(lv->level[j]) >> sh) - base] = 1;
    __syncthreads();
    int n_amb = assign_slots(s_flag, ncells, sh, (1 << (31 - sh)) - 1, fine_base, s_scan);
    if (t == 0) {
        pl.sh = sh; pl.cell_base = base; pl.ncells = ncells; pl.n_amb = n_amb;
        // fine table needs n_amb * 2^sh counters of 8 bytes
        bool fits = ((unsigned long long)n_amb << (sh + 3)) <= (1ull << fine_bytes_log2);
        pl.status = fits ? PLAN_HIST : PLAN_BSEARCH;
        *plan = pl;
    }
}

void papr_launch_plan_exact(const PaprDevLevels *lv, const PaprDevStats *merged, int fine_bytes_log2,
                            PaprPlan *plan, unsigned *fine_base, cudaStream_t s)
{
    papr_plan_exact_kernel<<<1, 1024, 0, s>>>(lv, merged, fine_bytes_log2, plan, fine_base);
}

__global__ void papr_zero_fine_kernel(const PaprPlan *plan, ulonglong2 *g_fine)
{
    const PaprPlan pl = *plan;
    if (!(pl.status & PLAN_HIST)) return;
    const size_t n2 = ((size_t)pl.n_amb << pl.sh) >> 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        g_fine[i] = make_ulonglong2(0ull, 0ull);
}

void papr_launch_zero_fine(const PaprPlan *plan, u64 *g_fine, int grid, cudaStream_t s)
{
    papr_zero_fine_kernel<<<grid, 512, 0, s>>>(plan, reinterpret_cast<ulonglong2 *>(g_fine));
}

// ------------------------------------------------------------------------------------------------
// resolve: cell histogram + fine table + exact thresholds -> level_count   (papr.c:147-151 restated)
// ------------------------------------------------------------------------------------------------
// one CTA per level (grid-stride): counts[j] = samples above the cell range + samples in the cells
// above cell(T_j) + the fine counters of the values of that cell that are > T_j.  A threshold outside
// the planned cells / windows => RES_MISS.
__global__ void __launch_bounds__(256) papr_count_kernel(const PaprPlan *plan, const unsigned *fine_base,
                                                         const PaprDevLevels *lv, const u64 *g_hist,
                                                         const u64 *g_fine, const u64 *g_over, u64 *counts,
                                                         u64 *status_word)
{
    __shared__ u64 s_red[256];
    const PaprPlan pl = *plan;
    const int L = lv->L;
    if (pl.status & PLAN_BSEARCH) return; // the generic kernel produced the counts
    for (int j = blockIdx.x; j < L; j += gridDim.x) {
        unsigned tb = __float_as_uint(lv->level[j]);
        int d = (int)(tb >> pl.sh) - pl.cell_base;
        bool in = (pl.status & PLAN_HIST) && (unsigned)d < (unsigned)pl.ncells;
        unsigned fb = in ? fine_base[d] : 0;
        if (!fb) {
            if (threadIdx.x == 0) { atomicOr(status_word, (u64)RES_MISS); counts[j] = 0; }
            continue;
        }
        const unsigned fmask = (1u << pl.sh) - 1u;
        const u64 *f = g_fine + (fb - 1u);
        u64 acc = 0;
        for (unsigned k = (tb & fmask) + 1 + threadIdx.x; k <= fmask; k += 256) acc += f[k];
        for (int c = d + 1 + threadIdx.x; c < pl.ncells; c += 256) acc += g_hist[c];
        s_red[threadIdx.x] = acc;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) counts[j] = s_red[0] + *g_over;
        __syncthreads();
    }
}

void papr_launch_resolve(const PaprPlan *plan, const unsigned *fine_base, const PaprDevLevels *lv,
                         const u64 *g_hist, const u64 *g_fine, const u64 *g_over, u64 *counts, u64 *status_word,
                         int grid, cudaStream_t s)
{
    papr_count_kernel<<<grid, 256, 0, s>>>(plan, fine_base, lv, g_hist, g_fine, g_over, counts, status_word);
}

// ------------------------------------------------------------------------------------------------
// peer-memory exchange (one process per GPU, windows mapped into every rank with cudaIpc): the three
// latency-sized exchanges of a sharded analysis, done inside the kernels that produce and consume the
// values instead of three collective launches in between.  papr_device.cuh: PaprXchg.
// ------------------------------------------------------------------------------------------------

#include "papr_xchg.cuh"

// fused mode, sharded: fold this GPU's presample triples, publish them, wait for the other ranks',
// add them up in rank order (identical on every rank) and plan from the global prediction
__global__ void __launch_bounds__(1024) papr_plan_pred_x_kernel(const double *cta_pre, int nctas, PaprTables tb,
                                                                float sigmas, float bias, int fine_slots, PaprPlan *plan,
                                                                unsigned *fine_base, PaprPeers pp, u64 seq, int ncells_max)
{
    __shared__ u64 s_mine[4];
    double pre[3];
    reduce_pre(cta_pre, nctas, pre);
    if (threadIdx.x == 0) {
        s_mine[0] = (u64)__double_as_longlong(pre[0]);
        s_mine[1] = (u64)__double_as_longlong(pre[1]);
        s_mine[2] = (u64)__double_as_longlong(pre[2]);
        s_mine[3] = 0;
    }
    __syncthreads();
    xchg_publish(pp, XK_PRE, offsetof(PaprXchgSlot, pre), s_mine, 4, seq);
    const bool ok = xchg_wait(pp, XK_PRE, seq);
    pre[0] = pre[1] = pre[2] = 0.0;
    for (int q = 0; q < pp.world; ++q) {
        const u64 *src = reinterpret_cast<const u64 *>(pp.win[pp.rank]->slot[q].pre);
        pre[0] += __longlong_as_double((long long)ld_volatile(src));
        pre[1] += __longlong_as_double((long long)ld_volatile(src + 1));
        pre[2] += __longlong_as_double((long long)ld_volatile(src + 2));
    }
    if (!ok) pre[2] = 0.0; // no plan: the result is reported as failed anyway
    plan_pred_body(pre, tb, sigmas, bias, fine_slots, plan, fine_base, ok ? 0 : 1, ncells_max);
}

void papr_launch_plan_pred_x(const double *cta_pre, int nctas, PaprTables t, float sigmas, float bias, int fine_slots,
                             PaprPlan *plan, unsigned *fine_base, PaprPeers pp, u64 seq, cudaStream_t s, int ncells_max)
{
    papr_plan_pred_x_kernel<<<1, 1024, 0, s>>>(cta_pre, nctas, t, sigmas, bias, fine_slots, plan, fine_base, pp, seq, ncells_max);
}

// sharded: CTA partials -> this shard's pass-1 state -> published; every rank's state collected and merged in rank
// order; avg, L and the level table of the WHOLE capture (papr_finalize.cuh: finalize_levels_x_body).  Stand-alone
// launch for shards that do not chain the sequential sum; otherwise the extra CTA of papr_xt_compose_kernel does it.
__global__ void __launch_bounds__(FIN_T) papr_finalize_levels_x_kernel(const PaprFinalizeXArgs a)
{
    finalize_levels_x_body(a);
}

void papr_launch_finalize_levels_x(const PaprFinalizeXArgs &a, cudaStream_t s)
{
    papr_finalize_levels_x_kernel<<<1, FIN_T, 0, s>>>(a);
}

// sharded: publish this shard's level counts (+ status word), collect every rank's, add them up (papr_xchg.cuh).
// Stand-alone launch for the recount / exact-redo paths; the first exchange of an analysis happens in the tail of
// papr_xt_epilogue_x_kernel.
__global__ void __launch_bounds__(1024) papr_counts_x_kernel(u64 *counts, const PaprDevLevels *lv, PaprPlan *plan,
                                                             PaprPeers pp, u64 seq)
{
    xchg_counts(counts, min(max(lv->L, 0), PAPR_MAX_LEVELS), plan, pp, seq);
}

void papr_launch_counts_x(u64 *counts, const PaprDevLevels *lv, PaprPlan *plan, PaprPeers pp, u64 seq, cudaStream_t s)
{
    papr_counts_x_kernel<<<1, 1024, 0, s>>>(counts, lv, plan, pp, seq);
}

// ------------------------------------------------------------------------------------------------
// generic CCDF kernel (any level table that is non-decreasing): per-sample binary search.
// Runs only when the exact plan does not fit the fine table (PLAN_BSEARCH) - e.g. PAPR > ~100 dB.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PAPR_THREADS, 1) papr_bsearch_kernel(const float *iq, u64 nsamples,
                                                                       const PaprPlan *plan,
                                                                       const PaprDevLevels *lv, u64 *bhist)
{
    __shared__ float s_level[PAPR_MAX_LEVELS];
    __shared__ unsigned s_h[PAPR_MAX_LEVELS + 1];
    if (!(plan->status & PLAN_BSEARCH)) return;
    const int L = lv->L;
    for (int i = threadIdx.x; i <= L; i += PAPR_THREADS) {
        s_h[i] = 0;
        if (i < L) s_level[i] = lv->level[i];
    }
    __syncthreads();
    const float2 *p = reinterpret_cast<const float2 *>(iq);
    for (u64 i = (u64)blockIdx.x * PAPR_THREADS + threadIdx.x; i < nsamples; i += (u64)gridDim.x * PAPR_THREADS) {
        float2 q = p[i];
        float v = power_of(q.x, q.y);
        int lo = 0, hi = L; // c(v) = #{j : level[j] < v}
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (s_level[mid] < v) lo = mid + 1; else hi = mid;
        }
        if (lo) atomicAdd(&s_h[lo], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= L; i += PAPR_THREADS)
        if (s_h[i]) atomicAdd(&bhist[i], (u64)s_h[i]);
}

void papr_launch_bsearch(const float *iq, u64 nsamples, const PaprPlan *plan, const PaprDevLevels *lv,
                         u64 *bhist, int grid, cudaStream_t s)
{
    papr_bsearch_kernel<<<grid, PAPR_THREADS, 0, s>>>(iq, nsamples, plan, lv, bhist);
}

// counts[j] = sum_{k > j} bhist[k]   (one thread; L <= 2048)
__global__ void papr_bsearch_counts_kernel(const PaprPlan *plan, const PaprDevLevels *lv, const u64 *bhist,
                                           u64 *counts)
{
    if (!(plan->status & PLAN_BSEARCH)) return;
    u64 above = 0;
    for (int j = lv->L - 1; j >= 0; --j) {
        above += bhist[j + 1];
        counts[j] = above;
    }
}

void papr_launch_bsearch_counts(const PaprPlan *plan, const PaprDevLevels *lv, const u64 *bhist, u64 *counts,
                                cudaStream_t s)
{
    papr_bsearch_counts_kernel<<<1, 1, 0, s>>>(plan, lv, bhist, counts);
}

// ------------------------------------------------------------------------------------------------
// synthetic capture generator, bit-identical to its C and numpy twins
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 mix64(u64 z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float siggen_component(u64 idx, unsigned c, u64 seed)
{
    int s = -262140;
#pragma unroll
    for (unsigned w = 0; w < 2; ++w) {
        u64 z = mix64((4 * idx + 2 * c + w + 1) * 0x9E3779B97F4A7C15ULL + seed);
        s += (int)((z & 0xffff) + ((z >> 16) & 0xffff) + ((z >> 32) & 0xffff) + (z >> 48));
    }
    return (float)s * 0x1p-19f;
}

__global__ void papr_siggen_kernel(float2 *iq, u64 first, u64 nsamples, u64 seed)
{
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < nsamples; k += (u64)gridDim.x * blockDim.x)
        iq[k] = make_float2(siggen_component(first + k, 0, seed), siggen_component(first + k, 1, seed));
}

void papr_launch_siggen(float *iq, u64 first, u64 nsamples, u64 seed, int grid, cudaStream_t s)
{
    papr_siggen_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<float2 *>(iq), first, nsamples, seed);
}

// first sample whose power is NaN (the sign of the reference's printed "nan" follows that sample)
__global__ void papr_find_nan_kernel(const float *iq, u64 nsamples, u64 first_index, u64 *out_idx)
{
    const float2 *p = reinterpret_cast<const float2 *>(iq);
    u64 best = ~0ull;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nsamples; i += (u64)gridDim.x * blockDim.x) {
        float2 q = p[i];
        float v = power_of(q.x, q.y);
        if (v != v) { best = first_index + i; break; }
    }
    if (best != ~0ull) atomicMin(out_idx, best);
}

void papr_launch_find_nan(const float *iq, u64 nsamples, u64 first_index, u64 *out_idx, int grid,
                          cudaStream_t s)
{
    papr_find_nan_kernel<<<grid, 256, 0, s>>>(iq, nsamples, first_index, out_idx);
}

// ------------------------------------------------------------------------------------------------
// exact emulation of the reference's SEQUENTIAL double sum (papr.c:104)        SURVEY.md §7.3
// ------------------------------------------------------------------------------------------------
// sum += (double)value, one sample after the other, rounds to nearest-even at every step; a parallel
// sum differs from it in the last bits.  Within one binade of the running sum (ulp u = 2^(k-52)) the
// state is an integer m = sum/u and adding v = (q + f)*u gives m' = m + q + r, where r = 1 iff f > 1/2,
// or f == 1/2 and m+q is odd - so the effect of a run of samples depends on the entry state only
// through its PARITY.  A run is therefore a pair (E0, E1): the total increment for an even / odd entry
// state.  Runs compose associatively (C.Ep = A.Ep + B.E[parity after A]), which makes the sequential
// sum a parallel reduction - exact as long as the running sum stays inside the binade, which the
// host guarantees per tile from (approximate) prefix sums and re-checks when it chains the tiles.
// Tiles that may cross a power of two are replayed on the host with real double adds.

// sums of the tiles (only used to place every tile in a binade; any summation order will do)
__global__ void __launch_bounds__(1024) papr_tilesum_kernel(const float *iq, u64 nsamples, double *tile_sum)
{
    __shared__ double s_w[32];
    const unsigned ntiles = (unsigned)((nsamples + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE);
    const float4 *p = reinterpret_cast<const float4 *>(iq);
    for (unsigned t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const u64 s0 = (u64)t * PAPR_SEQ_TILE;
        double acc = 0.0;
#pragma unroll 4
        for (int j = 0; j < PAPR_SEQ_TILE / 2 / 1024; ++j) {
            const u64 s = s0 + 2ull * (threadIdx.x + 1024u * j);
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s + 1 < nsamples) q = ldg_stream(p + (s >> 1));
            else if (s < nsamples) { float2 h = *reinterpret_cast<const float2 *>(iq + 2 * s); q.x = h.x; q.y = h.y; }
            acc += (double)power_of(q.x, q.y);
            acc += (double)power_of(q.z, q.w);
        }
        acc = warp_sum_fixed(acc);
        if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            double x = warp_sum_fixed(s_w[threadIdx.x]);
            if (threadIdx.x == 0) tile_sum[t] = x;
        }
        __syncthreads();
    }
}

void papr_launch_tilesum(const float *iq, u64 nsamples, double *tile_sum, int grid, cudaStream_t s)
{
    papr_tilesum_kernel<<<grid, 1024, 0, s>>>(iq, nsamples, tile_sum);
}

// A run = what a stretch of samples does to the running sum inside ONE binade [2^k, 2^(k+1)): the
// increment for an even and for an odd entry state.  The hardware computes both: an accumulator that
// starts at the bottom of the binade, 2^k (integer state 2^52: even) or 2^k + ulp (odd), has the ulp of
// the real running sum and the parity of the hypothesis, so its IEEE adds round exactly like the
// reference's `sum += value` does - ties to even included - for as long as it stays inside the binade.
// Two DADDs per sample instead of an integer emulation.  An accumulator that leaves the binade only
// grows from there, so the increment it reports is >= 2^k and the chain (seq_chain) rejects the run.
struct SeqRun { double a0, a1; }; // accumulators: 2^k (+ ulp) + everything added so far

__device__ __forceinline__ double seq_base(int k, int odd)
{
    return __longlong_as_double(((long long)(k + 1023) << 52) | (long long)odd);
}

__device__ __forceinline__ unsigned seq_lsb(double a) { return (unsigned)__double2loint(a) & 1u; }

// run A followed by run B (both started from the bases c0 / c1 of the same binade)
__device__ __forceinline__ SeqRun seq_compose(const SeqRun &a, const SeqRun &b, double c0, double c1)
{
    const double i0 = __dsub_rn(b.a0, c0), i1 = __dsub_rn(b.a1, c1); // B's increments: exact
    SeqRun c;
    c.a0 = __dadd_rn(a.a0, seq_lsb(a.a0) ? i1 : i0);
    c.a1 = __dadd_rn(a.a1, seq_lsb(a.a1) ? i1 : i0);
    return c;
}

// One CTA of SEQ_T threads per tile of PAPR_SEQ_TILE samples; thread t owns the SEQ_PER consecutive
// samples t*SEQ_PER.. of the tile and walks them in file order straight from global memory (16-byte
// loads, 4 in flight; the two halves of every 32-byte sector are consumed back to back), then the
// threads' runs are composed in thread (= file) order: shuffles inside a warp, one pass over the
// warps.  Several CTAs per SM are resident, each in a different phase, which hides the latency.
#define SEQ_T 256
#define SEQ_PER (PAPR_SEQ_TILE / SEQ_T) // 128 samples = 1 KiB per thread
__global__ void __launch_bounds__(SEQ_T) papr_seqsum_kernel(const float *iq, u64 nsamples, const short *tile_code,
                                                           double *tile_run)
{
    __shared__ SeqRun s_w[SEQ_T / 32];
    const unsigned ntiles = (unsigned)((nsamples + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int k = tile_code[tile];
        if (k >= PAPR_SEQ_ZERO) continue; // zero or dirty tile: nothing to do here (uniform per CTA)
        const double c0 = seq_base(k, 0), c1 = seq_base(k, 1);
        const u64 s0 = (u64)tile * PAPR_SEQ_TILE + (u64)t * SEQ_PER;
        const float4 *p = reinterpret_cast<const float4 *>(iq + 2 * s0);
        SeqRun r;
        r.a0 = c0; r.a1 = c1;
        if (s0 + SEQ_PER <= nsamples) {
#pragma unroll 1
            for (int j = 0; j < SEQ_PER / 2; j += 4) {
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) q[u] = __ldg(p + j + u);
#pragma unroll
                for (int u = 0; u < 4; ++u) { // papr.c:103-104, twice: once per hypothesis
                    const double v0 = (double)power_of(q[u].x, q[u].y), v1 = (double)power_of(q[u].z, q[u].w);
                    r.a0 = __dadd_rn(__dadd_rn(r.a0, v0), v1);
                    r.a1 = __dadd_rn(__dadd_rn(r.a1, v0), v1);
                }
            }
        } else { // ragged end of the capture
            for (u64 s = s0; s < nsamples && s < s0 + SEQ_PER; ++s) {
                const float2 h = *reinterpret_cast<const float2 *>(iq + 2 * s);
                const double v = (double)power_of(h.x, h.y);
                r.a0 = __dadd_rn(r.a0, v);
                r.a1 = __dadd_rn(r.a1, v);
            }
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { // ordered: lane l (multiple of 2o) <- l then l+o
            SeqRun nb;
            nb.a0 = __shfl_down_sync(FULL, r.a0, o);
            nb.a1 = __shfl_down_sync(FULL, r.a1, o);
            r = seq_compose(r, nb, c0, c1); // only the lanes that are multiples of 2o keep a meaningful value
        }
        if (lane == 0) s_w[warp] = r;
        __syncthreads();
        if (t == 0) {
            SeqRun acc = s_w[0];
            for (int w = 1; w < SEQ_T / 32; ++w) acc = seq_compose(acc, s_w[w], c0, c1);
            tile_run[2 * tile] = __dsub_rn(acc.a0, c0);     // increment for an even entry state
            tile_run[2 * tile + 1] = __dsub_rn(acc.a1, c1); // ... and for an odd one
        }
        __syncthreads();
    }
}

int papr_seqsum_configure(void) { return 0; } // no dynamic shared memory any more

void papr_launch_seqsum(const float *iq, u64 nsamples, const short *tile_code, void *tile_run, int grid,
                        cudaStream_t s)
{
    papr_seqsum_kernel<<<grid, SEQ_T, 0, s>>>(iq, nsamples, tile_code, reinterpret_cast<double *>(tile_run));
}
