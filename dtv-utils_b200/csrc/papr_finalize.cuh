// papr_finalize.cuh — the pass-1 follow-ups shared by papr_kernels.cu (stand-alone launches) and papr_exact.cu
// (the sharded chained path runs the statistics exchange as an extra CTA of the compose kernel): fold of the CTA
// partials, merge of the shards' states + the reference's scalar epilogue (papr.c:131-141) on the device, and
// the sharded variant with the exchange of the pass-1 states inside.
#pragma once
#include "papr_scan_common.cuh"
#include "papr_xchg.cuh"

#define FIN_T 256
// fixed-order reduction of the CTA partials (sum) and (value desc, index asc) selection; result in
// thread 0 of the calling CTA
static __device__ void reduce_partials(const PaprCtaPartial *wp, int nctas, u64 n, PaprDevStats *out)
{
    __shared__ double s_sum[FIN_T];
    __shared__ int s_val[PAPR_NTRACK][FIN_T];
    __shared__ u64 s_idx[PAPR_NTRACK][FIN_T];
    const int t = threadIdx.x;
    const bool on = t < FIN_T; // (a larger CTA may call this: the threads beyond FIN_T only keep the barriers)
    double sum = 0.0;
    int val[PAPR_NTRACK];
    u64 idx[PAPR_NTRACK];
    for (int k = 0; k < PAPR_NTRACK; ++k) { val[k] = 0; idx[k] = 0; }
    for (int i = t; on && i < nctas; i += FIN_T) {
        sum += wp[i].sum;
        for (int k = 0; k < PAPR_NTRACK; ++k)
            if (wp[i].val[k] > 0 && better(wp[i].val[k], wp[i].idx[k], val[k], idx[k])) {
                val[k] = wp[i].val[k];
                idx[k] = wp[i].idx[k];
            }
    }
    if (on) {
        s_sum[t] = sum;
        for (int k = 0; k < PAPR_NTRACK; ++k) { s_val[k][t] = val[k]; s_idx[k][t] = idx[k]; }
    }
    __syncthreads();
    for (int o = FIN_T / 2; o > 0; o >>= 1) {
        if (t < o) {
            s_sum[t] += s_sum[t + o];
            for (int k = 0; k < PAPR_NTRACK; ++k)
                if (s_val[k][t + o] > 0 && better(s_val[k][t + o], s_idx[k][t + o], s_val[k][t], s_idx[k][t])) {
                    s_val[k][t] = s_val[k][t + o];
                    s_idx[k][t] = s_idx[k][t + o];
                }
        }
        __syncthreads();
    }
    if (t == 0) {
        out->sum = s_sum[0];
        out->n = n;
        for (int k = 0; k < PAPR_NTRACK; ++k) { out->val[k] = s_val[k][0]; out->idx[k] = s_idx[k][0]; }
        out->flags = isfinite(s_sum[0]) ? 0u : PAPR_FLAG_NONFINITE;
    }
}

// Merge the shards' pass-1 states in rank (= index) order and evaluate the reference's scalar
// epilogue on the device:  avg = sum/offset (papr.c:131), ratio = peak/avg, L, and
// level[j] = (float)(pow(10, x_j) * avg) (papr.c:139 / 170).  pow(10, x_j) and the least ratio for
// which the reference's loops reach level j are tabulated once with the HOST libm, so only IEEE
// double multiply/divide/compare and one rounding to float happen here - bit-identical to the host.
static __device__ void merge_and_levels(const PaprDevStats *parts, int nparts, const PaprTables &tb, int graph,
                                 PaprDevStats *merged, PaprDevLevels *lv, u64 *status_word,
                                 size_t stride_bytes = sizeof(PaprDevStats), const PaprChainList *chain = nullptr,
                                 int *chain_report = nullptr, double bias = 1.0 /* test hook: scales the sum the levels come from */)
{
    __shared__ int s_L;
    __shared__ double s_avg;
    if (threadIdx.x == 0) {
        PaprDevStats m = parts[0];
        for (int p = 1; p < nparts; ++p) {
            const PaprDevStats q = *reinterpret_cast<const PaprDevStats *>(reinterpret_cast<const char *>(parts) + p * stride_bytes);
            m.sum += q.sum;
            m.n += q.n;
            for (int k = 0; k < PAPR_NTRACK; ++k)
                if (q.val[k] > m.val[k]) { m.val[k] = q.val[k]; m.idx[k] = q.idx[k]; } // strict: earlier shard wins ties
            m.flags |= q.flags;
        }
        if (!isfinite(m.sum)) m.flags |= PAPR_FLAG_NONFINITE;
        if (chain) { // the sequential sum chained over all shards on the device replaces the approximate one
            const int st = chain->status;
            if (st == XT_OK) m.sum = chain->exact;
            chain_report[0] = st;
            chain_report[1] = chain->why;
        }
        *merged = m;
        double avg = __ddiv_rn(__dmul_rn(m.sum, bias), (double)(long long)m.n);
        double ratio = __ddiv_rn((double)__int_as_float(m.val[TR_PEAK]), avg);
        int lo = 0, hi = tb.nlevels_max; // number of j with ratio >= ratio_min[j] (non-decreasing table)
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (ratio >= tb.ratio_min[mid]) lo = mid + 1; else hi = mid;
        }
        lv->avg = avg;
        lv->ratio = ratio;
        lv->L = lo;
        lv->graph = graph;
        *status_word = 0; // RES_* bits of the resolve that follows
        s_L = lo;
        s_avg = avg;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < s_L; j += blockDim.x)
        lv->level[j] = __double2float_rn(__dmul_rn(tb.pow10[j], s_avg));
}


// All threads of ONE CTA (>= FIN_T threads).
static __device__ void finalize_levels_x_body(const PaprFinalizeXArgs &a)
{
    reduce_partials(a.wp, a.nctas, a.n, a.local);
    __threadfence();
    __syncthreads(); // thread 0's *local is visible to the CTA
    xchg_publish(a.pp, XK_STATS, offsetof(PaprXchgSlot, stats), reinterpret_cast<const u64 *>(a.local),
                 (int)(sizeof(PaprDevStats) / 8), a.seq);
    const bool ok = xchg_wait(a.pp, XK_STATS, a.seq);
    if (!ok && threadIdx.x == 0) a.plan->pad = 1;
    __shared__ PaprDevStats s_parts[PAPR_XCHG_MAX_RANKS];
    for (int i = threadIdx.x; i < a.pp.world * (int)(sizeof(PaprDevStats) / 8); i += blockDim.x) {
        const int q = i / (int)(sizeof(PaprDevStats) / 8), w = i % (int)(sizeof(PaprDevStats) / 8);
        reinterpret_cast<u64 *>(&s_parts[q])[w] = ld_volatile(reinterpret_cast<const u64 *>(&a.pp.win[a.pp.rank]->slot[q].stats) + w);
    }
    __syncthreads();
    // (chained shards: these levels come from the fixed-order sums; the host checks them against the chained sum's)
    merge_and_levels(s_parts, a.pp.world, a.tb, a.graph, a.merged, a.lv, a.status_word, sizeof(PaprDevStats), nullptr, nullptr, a.bias);
}
