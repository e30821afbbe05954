// papr_exact.cu — the fused scan of a device-resident shard, fed by TMA, with the reference's
// SEQUENTIAL double sum (drmpeg/dtv-utils papr.c:104, `sum += value` in file order, rounding to
// nearest-even at every add) emulated bit for bit inside the same sweep.
//
//   papr_scan_tma_kernel   papr.c:100-129 + 143-153 in one pass, like papr_scan_kernel<1,1>, but the samples
//                          arrive through 2-D TMA boxes (16 rows x 128 B = one 2 KiB warp batch, 128-byte
//                          swizzle) so that every lane reads a run of 8 CONSECUTIVE samples without bank
//                          conflicts, and every warp owns whole tiles of 16 consecutive batches.
//                          Per lane and batch two DADD accumulators (even / odd entry state, anchored
//                          at the bottom of the predicted binade of the running sum) do the rounding
//                          exactly as the reference's adds do; two ballots carry the parity from lane to
//                          lane; per-lane totals are folded once per tile.
//   papr_xt_compose_kernel 32 tile runs -> one super-tile run (ordered, associative composition)
//   the chain              one CTA (CTA 0 of papr_xt_epilogue_kernel; papr_xt_chain_x_kernel when sharded):
//                          approximate prefix sums locate every binade crossing of the running sum down to
//                          one sample; the stretches between crossings are composed in parallel; what
//                          remains is a list of < 100 items applied in order.
//   papr_xt_epilogue_kernel (single shard) the chain side by side with finalize + levels + level counts
// Everything is verified while it is applied (state inside the assumed binade before and after each run);
// a failed check only ever sends the analysis to the slower two-sweep emulation, never to a wrong sum.
#include <cuda.h>
#include <cstdio>

#include "papr_scan_common.cuh"
#include "papr_xchg.cuh"
#include "papr_finalize.cuh"

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double xt_base(int k, unsigned odd) // 2^k, or 2^k + one ulp
{
    return __longlong_as_double(((long long)(k + 1023) << 52) | (long long)odd);
}

__device__ __forceinline__ unsigned xt_lsb(double a) { return (unsigned)__double2loint(a) & 1u; }

// floor(log2(x)) of a positive normal double; -4000 for zero / denormals / negatives, +4000 for Inf / NaN
__device__ __forceinline__ int xt_expo(double x)
{
    const long long b = __double_as_longlong(x);
    const int e = (int)((b >> 52) & 0x7ff);
    if (b <= 0 || e == 0) return -4000;
    if (e == 0x7ff) return 4000;
    return e - 1023;
}

__device__ __forceinline__ PaprTileRun xt_compose(const PaprTileRun &a, const PaprTileRun &b, int k)
{
    // parity of the state after A: bit 0 of (base + increment); exact while A stays inside the binade, and
    // an A that left it keeps e >= 2^k whatever is added (increments are >= 0), which the walk rejects
    const unsigned p0 = xt_lsb(__dadd_rn(xt_base(k, 0), a.e0)), p1 = xt_lsb(__dadd_rn(xt_base(k, 1), a.e1));
    PaprTileRun c;
    c.e0 = __dadd_rn(a.e0, p0 ? b.e1 : b.e0);
    c.e1 = __dadd_rn(a.e1, p1 ? b.e1 : b.e0);
    return c;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase)
{
    unsigned done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
}

// one 4 KiB box: rows [row, row+32) x 128 B of the capture -> swizzled shared memory, completion on `bar`.
// Called by the whole (converged) warp; one elected lane arms the barrier and issues the copy.  Every byte of the
// capture is used once: the L2 policy `pol` (evict-first) keeps the stream from pushing the fine table, the
// histogram and the tile records out of L2 (they are what the sweep writes back to DRAM otherwise).
__device__ __forceinline__ void tma_batch(unsigned dst, const CUtensorMap *tm, int row, unsigned bar, unsigned long long pol)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "elect.sync _|p, 0xffffffff;\n\t"
                 "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4096;\n\t"
                 "@p cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%1], [%2, {%3, %4}], [%0], %5;\n\t"
                 "}" ::"r"(bar), "r"(dst), "l"(tm), "r"(0), "r"(row), "l"(pol) : "memory");
}

__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}

// byte offset of sample i (0..511) of a batch inside its swizzled 4 KiB box
__device__ __forceinline__ unsigned xt_sample_off(unsigned i)
{
    const unsigned row = i >> 4, chunk = (i >> 1) & 7;
    return row * 128u + ((chunk ^ (row & 7u)) << 4) + (i & 1u) * 8u;
}

// Candidate binades of the running sum while samples [g0, g1) of the whole capture are added, from the
// presample's mean and its uncertainty.  Returns the number of candidates (0 = cannot tell), first in *k_lo.
__device__ __forceinline__ int xt_candidates(double avg, double window, double wcv, unsigned long long g0,
                                             unsigned long long g1, int *k_lo)
{
    if (!(avg > 0.0) || g0 == 0) return 0;
    const double w = window + wcv * rsqrt((double)g0 * (1.0 / PAPR_BATCH_SAMPLES));
    const double lo = avg * (double)g0 * (1.0 - w), hi = avg * (double)g1 * (1.0 + w);
    const int a = xt_expo(lo), b = xt_expo(hi);
    if (a <= -4000 || b >= 4000 || b - a >= XT_MAX_CAND) return 0;
    *k_lo = a;
    return b - a + 1;
}

// ------------------------------------------------------------------------------------------------
// the fused scan
// ------------------------------------------------------------------------------------------------
#ifndef XT_THREADS
#define XT_THREADS 640 // 20 warps x 4 KiB in flight; ~100 registers per thread (24 warps: more spills than gain, profiles/README.md)
#endif
#define XT_WARPS (XT_THREADS / 32)
#define XT_RING_BYTES (XT_WARPS * 4096)
#define XT_HIST_BYTES (8 * (XT_NCELLS + 2))
#define XT_SMEM_BYTES (XT_HIST_BYTES + 1024 + XT_RING_BYTES)

// entry parity of this lane's run given the exit parities of the lower lanes (b0 / b1: ballots of the exit
// parity for an even / odd entry) and the parity P with which the batch is entered
struct XtLink {
    unsigned flip, fixed, base; // parity = (fixed ? base : P) ^ flip
};

__device__ __forceinline__ XtLink xt_link(unsigned b0, unsigned b1, unsigned lt)
{
    const unsigned c = ~(b0 ^ b1); // lanes whose exit parity does not depend on how they were entered
    const unsigned ng = b0 & ~b1;  // the others either pass the parity on (0,1) or invert it (1,0): these invert
    const unsigned cm = c & lt;
    const unsigned upto = cm ? (0xffffffffu >> __clz(cm)) : 0u; // lanes 0..j, j = nearest such lane below
    XtLink l;
    l.fixed = cm != 0;
    l.base = (b0 & (upto ^ (upto >> 1))) != 0; // exit parity of lane j
    l.flip = __popc(ng & lt & ~upto) & 1u;
    return l;
}

// first sample of the batch (position inside it) whose tracked value equals the warp-wide maximum w
template <int T>
__device__ __forceinline__ unsigned xt_locate(const float4 (&r)[XT_RUN / 2], int w, int lane)
{
    unsigned pos = 0xffffffffu;
#pragma unroll
    for (int u = XT_RUN / 2 - 1; u >= 0; --u) {
        float v0 = power_of(r[u].x, r[u].y), v1 = power_of(r[u].z, r[u].w), a, b;
        track_vals<T>(r[u], v0, v1, a, b);
        const unsigned p = (unsigned)XT_RUN * (unsigned)lane + 2u * (unsigned)u;
        if (__float_as_int(b) == w) pos = p + 1;
        if (__float_as_int(a) == w) pos = p;
    }
    return __reduce_min_sync(FULL, pos);
}

// smallest power a sample needs to have a chance against any of the five records (see the sweep)
__device__ __forceinline__ float xt_threshold(const int (&val)[PAPR_NTRACK])
{
    float m = __int_as_float(val[TR_RE_POS]);
    m = fminf(m, __int_as_float(val[TR_RE_NEG]));
    m = fminf(m, __int_as_float(val[TR_IM_POS]));
    m = fminf(m, __int_as_float(val[TR_IM_NEG])); // magnitudes (>= 0); 0 = no sample of that sign yet: every batch passes
    return __fmul_rd(m, m);
}

__global__ void __launch_bounds__(XT_THREADS, 1) papr_scan_tma_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                      const PaprScanArgs a, const PaprExactArgs x)
{
    ScanState<true, true> st; // the CCDF fields; the extremes live in run_val[] + shared memory (rarely written)
    __shared__ __align__(8) unsigned long long s_bar[XT_WARPS];
    __shared__ unsigned s_pos[PAPR_NTRACK][XT_WARPS]; // sample offset (inside this launch) of the warp's current first occurrences
    __shared__ int s_val[PAPR_NTRACK][XT_WARPS];      // ... and their values; 0 = nothing new from this warp
    __shared__ PaprTileRun s_mt[XT_WARPS][XT_MAX_CAND];
    unsigned *s_hist = reinterpret_cast<unsigned *>(scan_smem);
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0); // warp-uniform for the compiler: TMA operands stay in uniform registers
    const unsigned lt = (1u << lane) - 1u;

    const PaprPlan pl = *a.plan;
    const bool do_hist = (pl.status & PLAN_HIST) != 0;
    {
        unsigned *s_fb = s_hist + (XT_NCELLS + 2);
        st.g_fine = a.g_fine;
        asm volatile("{\n\t.reg .u64 t;\n\tcvta.to.shared.u64 t, %1;\n\tcvt.u32.u64 %0, t;\n\t}"
                     : "=r"(st.smem_slot0) : "l"(s_hist));
        st.sh = pl.sh; // (always PAPR_SH_MIN here: this sweep only follows papr_plan_pred*, the shift is compiled in)
        st.neg_base1 = 1 - (do_hist ? pl.cell_base : 0x7fffffff); // no valid plan: every sample lands in slot 0
        st.ncells = do_hist ? min(pl.ncells, XT_NCELLS) : 0;
        st.fmask = (1u << pl.sh) - 1u;
        for (int i = threadIdx.x; i < st.ncells + 2; i += XT_THREADS) {
            s_hist[i] = 0;
            s_fb[i] = (i >= 1 && i <= st.ncells) ? a.fine_base[i - 1] : 0u;
        }
    }
#pragma unroll
    for (int t = 0; t < PAPR_NTRACK; ++t) {
        st.run_val[t] = a.wp[blockIdx.x].val[t]; // carried over from earlier launches of this shard
        if (lane == 0) { s_val[t][warp] = 0; s_pos[t][warp] = 0xffffffffu; }
    }
    const unsigned ring = (smem_u32(scan_smem) + XT_HIST_BYTES + 1023u) & ~1023u;
    const unsigned my = ring + (unsigned)warp * 4096u, bar = smem_u32(&s_bar[warp]);
    if (lane == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // this lane's run: samples 16*lane .. 16*lane+15 of the batch = one 128-byte row; its eight 16-byte pieces
    // sit at off0 ^ (j << 4) (128-byte swizzle: chunk index XOR row % 8)
    const unsigned off0 = my + (unsigned)lane * 128u + (((unsigned)lane & 7u) << 4);

    const u64 n = a.nsamples;
    const unsigned nbatch = (unsigned)((n + XT_BATCH_SAMPLES - 1) / XT_BATCH_SAMPLES);
    const unsigned ntiles = (nbatch + XT_TILE_BATCHES - 1) / XT_TILE_BATCHES;
    // the tensor map covers the full 128-byte rows; the < 16 samples past the last one are patched into this batch
    const unsigned tail_n = (unsigned)(n & 15), tail_batch = tail_n ? (unsigned)(n / XT_BATCH_SAMPLES) : 0xffffffffu;
    const unsigned tstride = gridDim.x * XT_WARPS;
    unsigned phase = 0;
    double wsum = 0.0; // lane 0: approximate sum of this warp's tiles (what the tree sum used to be)
    float thr = xt_threshold(st.run_val);

    unsigned tile = blockIdx.x * XT_WARPS + warp;
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    if (tile < ntiles) tma_batch(my, &tmap, (int)(tile * XT_TILE_BATCHES * 32), bar, pol);
    for (; tile < ntiles; tile += tstride) {
        // ---- which binade(s) will the running sum be in while this tile is added?
        const u64 g0 = x.g_first + (u64)tile * XT_TILE_SAMPLES;
        int k_lo = 0;
        int nc = xt_candidates(pl.avg_pred, (double)pl.window, pl.wcv, g0, g0 + XT_TILE_SAMPLES, &k_lo);
        const bool literal = x.literal_tile0 && g0 == 0;
        unsigned slot = 0;
        if (nc > 1) {
            if (lane == 0) slot = atomicAdd(x.multi_count, 1u);
            slot = __shfl_sync(FULL, slot, 0);
            if (slot >= x.multi_cap) nc = 0; // log full: the chain falls back if this tile matters
        }
        // nc == 1: this lane's share of the tile's increment for an even (A0) / odd (A1) tile entry;
        // otherwise A0 is the approximate sum (lane 0: multi / literal, every lane: unknown)
        double A0 = 0.0, A1 = 0.0;
        unsigned P = 2u; // bit 0 / bit 1: parity of the running state under the two hypotheses
        if (lane == 0) // multi tiles: the tile's run per candidate, composed batch by batch (rare: shared memory)
            for (int cd = 0; cd < XT_MAX_CAND; ++cd) s_mt[warp][cd].e0 = s_mt[warp][cd].e1 = 0.0;
        const unsigned b_first = tile * XT_TILE_BATCHES;
        const unsigned nb = min((unsigned)XT_TILE_BATCHES, nbatch - b_first);
        for (unsigned b = 0; b < nb; ++b) {
            mbar_wait(bar, phase);
            phase ^= 1;
            if (b_first + b == tail_batch) {
                if ((unsigned)lane < tail_n) {
                    const u64 s = (n & ~15ull) + lane;
                    const float2 h = *reinterpret_cast<const float2 *>(a.iq + 2 * s);
                    asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(my + xt_sample_off((unsigned)(s % XT_BATCH_SAMPLES))), "f"(h.x), "f"(h.y) : "memory");
                }
                __syncwarp();
            }
            float4 r[XT_RUN / 2];
#pragma unroll
            for (int u = 0; u < XT_RUN / 2; ++u) r[u] = lds128(off0 ^ (16u * u));
            if (literal) { // papr.c:103-104 as written, from a running sum of exactly 0
                if (lane == 0) {
                    for (unsigned i = 0; i < XT_BATCH_SAMPLES; ++i) {
                        float2 h;
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(h.x), "=f"(h.y) : "r"(my + xt_sample_off(i)));
                        A0 = __dadd_rn(A0, (double)power_of(h.x, h.y));
                    }
                }
            }

            // ---- power and extremes (papr.c:103,105-126).  One conservative test covers all five trackers: a sample
            //      can raise a component record r only if I^2 or Q^2 >= r^2, hence only if its power >= thr = the
            //      smallest squared component record (rounded down) - and the peak record is never below thr.  The
            //      chance of a batch passing falls like 1/(samples seen), the same as that of a real update.
            float v[XT_RUN];
            {
                float vm = 0.f;
#pragma unroll
                for (int u = 0; u < XT_RUN / 2; ++u) {
                    v[2 * u] = power_of(r[u].x, r[u].y);
                    v[2 * u + 1] = power_of(r[u].z, r[u].w);
                    vm = fmaxf(vm, fmaxf(v[2 * u], v[2 * u + 1])); // (drops NaN operands, like the reference's compares)
                }
                if (__any_sync(FULL, vm >= thr)) { // warp-uniform and rare after the first few batches
                    float bm0 = 0.f, bm1 = 0.f, bm2 = 0.f, bm3 = 0.f, bm4 = 0.f;
#pragma unroll
                    for (int u = 0; u < XT_RUN / 2; ++u) {
                        const float4 q = r[u];
                        bm0 = fmaxf(bm0, fmaxf(v[2 * u], v[2 * u + 1]));
                        bm1 = fmaxf(bm1, fmaxf(q.x, q.z));
                        bm2 = fmaxf(bm2, fmaxf(-q.x, -q.z));
                        bm3 = fmaxf(bm3, fmaxf(q.y, q.w));
                        bm4 = fmaxf(bm4, fmaxf(-q.y, -q.w));
                    }
                    const unsigned batch_off = (b_first + b) * XT_BATCH_SAMPLES;
                    const float bm[PAPR_NTRACK] = {bm0, bm1, bm2, bm3, bm4};
#pragma unroll
                    for (int t = 0; t < PAPR_NTRACK; ++t) {
                        const int w = __reduce_max_sync(FULL, __float_as_int(bm[t]));
                        if (w > st.run_val[t]) {
                            unsigned pos;
                            if (t == TR_PEAK) pos = xt_locate<TR_PEAK>(r, w, lane);
                            else if (t == TR_RE_POS) pos = xt_locate<TR_RE_POS>(r, w, lane);
                            else if (t == TR_RE_NEG) pos = xt_locate<TR_RE_NEG>(r, w, lane);
                            else if (t == TR_IM_POS) pos = xt_locate<TR_IM_POS>(r, w, lane);
                            else pos = xt_locate<TR_IM_NEG>(r, w, lane);
                            st.run_val[t] = w;
                            if (lane == 0) { s_val[t][warp] = w; s_pos[t][warp] = batch_off + pos; }
                        }
                    }
                    thr = xt_threshold(st.run_val);
                }
            }
            // the batch is in registers: fetch the next one of this tile, or the first one of this warp's next tile
            __syncwarp();
            {
                const unsigned nxt = b + 1 < nb ? b_first + b + 1 : (tile + tstride) * XT_TILE_BATCHES;
                if (nxt < nbatch) tma_batch(my, &tmap, (int)(nxt * 32u), bar, pol);
            }

            // ---- CCDF cells (papr.c:147-151)
#pragma unroll
            for (int u = 0; u < XT_RUN / 2; ++u) hist_pair<true, true, 4 * (XT_NCELLS + 2), PAPR_SH_MIN>(st, v[2 * u], v[2 * u + 1]);

            // ---- papr.c:104
            if (nc == 1) {
                const double c0 = xt_base(k_lo, 0), c1 = xt_base(k_lo, 1);
                double a0 = c0, a1 = c1;
#pragma unroll
                for (int j = 0; j < XT_RUN; ++j) {
                    const double d = (double)v[j];
                    a0 = __dadd_rn(a0, d);
                    a1 = __dadd_rn(a1, d);
                }
                const unsigned x0 = xt_lsb(a0), x1 = xt_lsb(a1);
                const XtLink l = xt_link(__ballot_sync(FULL, x0), __ballot_sync(FULL, x1), lt);
                const double i0 = __dsub_rn(a0, c0), i1 = __dsub_rn(a1, c1);
                const unsigned p0 = (l.fixed ? l.base : (P & 1u)) ^ l.flip, p1 = (l.fixed ? l.base : (P >> 1)) ^ l.flip;
                A0 = __dadd_rn(A0, p0 ? i1 : i0);
                A1 = __dadd_rn(A1, p1 ? i1 : i0);
                // the parity with which the next batch is entered = lane 31's exit parity under each hypothesis
                P = __shfl_sync(FULL, (p0 ? x1 : x0) | ((p1 ? x1 : x0) << 1), 31);
            } else if (nc > 1) { // one run per candidate binade and per batch, for the chain to choose from
                for (int cd = 0; cd < nc; ++cd) {
                    const double d0 = xt_base(k_lo + cd, 0), d1 = xt_base(k_lo + cd, 1);
                    double a0 = d0, a1 = d1;
#pragma unroll
                    for (int j = 0; j < XT_RUN; ++j) {
                        const double d = (double)v[j];
                        a0 = __dadd_rn(a0, d);
                        a1 = __dadd_rn(a1, d);
                    }
                    const XtLink l = xt_link(__ballot_sync(FULL, xt_lsb(a0)), __ballot_sync(FULL, xt_lsb(a1)), lt);
                    const double i0 = __dsub_rn(a0, d0), i1 = __dsub_rn(a1, d1);
                    const unsigned p0 = (l.fixed ? l.base : 0u) ^ l.flip, p1 = (l.fixed ? l.base : 1u) ^ l.flip;
                    PaprTileRun br;
                    br.e0 = warp_sum_fixed(p0 ? i1 : i0);
                    br.e1 = warp_sum_fixed(p1 ? i1 : i0);
                    if (lane == 0) {
                        x.multi[((size_t)slot * XT_MAX_CAND + cd) * XT_TILE_BATCHES + b] = br;
                        if (cd == 0) A0 = __dadd_rn(A0, br.e0);
                        s_mt[warp][cd] = xt_compose(s_mt[warp][cd], br, k_lo + cd);
                    }
                }
            } else if (!literal) {
#pragma unroll
                for (int j = 0; j < XT_RUN; ++j) A0 = __dadd_rn(A0, (double)v[j]);
            }
        }

        // ---- the tile's record
        PaprTileRun tr;
        int code;
        if (nc == 1) {
            tr.e0 = warp_sum_fixed(A0);
            tr.e1 = warp_sum_fixed(A1);
            code = (k_lo + XT_KBIAS) | (1 << 12);
        } else if (nc > 1) {
            tr.e0 = A0; tr.e1 = 0.0;
            code = (k_lo + XT_KBIAS) | (nc << 12) | (int)(slot << 16);
            if (lane == 0) {
                PaprTileRun z;
                z.e0 = z.e1 = 0.0;
                for (int cd = 0; cd < nc; ++cd)
                    for (unsigned b = nb; b < XT_TILE_BATCHES; ++b)
                        x.multi[((size_t)slot * XT_MAX_CAND + cd) * XT_TILE_BATCHES + b] = z;
                for (int cd = 0; cd < nc; ++cd) x.multi_tile[(size_t)slot * XT_MAX_CAND + cd] = s_mt[warp][cd];
            }
        } else if (literal) {
            tr.e0 = tr.e1 = A0;
            code = XT_CODE_LITERAL;
        } else {
            tr.e0 = warp_sum_fixed(A0); tr.e1 = 0.0;
            code = 0;
        }
        if (lane == 0) {
            wsum = __dadd_rn(wsum, tr.e0);
            x.tile_run[x.tile_base + tile] = tr;
            x.tile_code[x.tile_base + tile] = code;
        }
    }

    // ---- CTA-level fold of the warps' states (fixed order), then into the CTA's persistent partial
    {
        __shared__ double s_wsum[XT_WARPS];
        if (lane == 0) s_wsum[warp] = wsum;
        __syncthreads();
        if (warp == 0) {
            double xs = warp_sum_fixed(lane < XT_WARPS ? s_wsum[lane] : 0.0);
            PaprCtaPartial *w = a.wp + blockIdx.x;
            if (lane == 0) w->sum += xs;
#pragma unroll
            for (int t = 0; t < PAPR_NTRACK; ++t) {
                int vv = lane < XT_WARPS ? s_val[t][lane] : 0;
                unsigned pos = lane < XT_WARPS ? s_pos[t][lane] : 0xffffffffu;
                const int vmax = __reduce_max_sync(FULL, vv);
                const unsigned pmin = __reduce_min_sync(FULL, vv == vmax ? pos : 0xffffffffu);
                if (lane == 0 && vmax > w->val[t]) { // strictly greater than what earlier launches left: first occurrence kept
                    w->val[t] = vmax;
                    w->idx[t] = a.first_index + pmin;
                }
            }
        }
    }
    if (do_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < st.ncells; i += XT_THREADS) {
            unsigned c = s_hist[i + 1];
            if (c) atomicAdd(&a.g_hist[i], (u64)c);
        }
        if (threadIdx.x == 0 && s_hist[st.ncells + 1]) atomicAdd(a.g_over, (u64)s_hist[st.ncells + 1]);
    }
}

int papr_scan_tma_configure(void)
{
    return cudaFuncSetAttribute(papr_scan_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XT_SMEM_BYTES) == cudaSuccess ? 0 : -1;
}

void papr_launch_scan_tma(const void *tmap, int grid, const PaprScanArgs &a, const PaprExactArgs &x, cudaStream_t s)
{
    papr_scan_tma_kernel<<<grid, XT_THREADS, XT_SMEM_BYTES, s>>>(*reinterpret_cast<const CUtensorMap *>(tmap), a, x);
}

// ------------------------------------------------------------------------------------------------
// 32 tile runs -> one super-tile record (one warp per super-tile); multi tiles: 16 batch runs -> one tile
// run per candidate binade
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ PaprTileRun xt_shfl_down(const PaprTileRun &r, int o)
{
    PaprTileRun n;
    n.e0 = __shfl_down_sync(FULL, r.e0, o);
    n.e1 = __shfl_down_sync(FULL, r.e1, o);
    return n;
}

// ordered composition of the 32 lanes' runs (identity = {0, 0}); lane 0 holds the result
__device__ __forceinline__ PaprTileRun xt_warp_compose(PaprTileRun r, int k)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) r = xt_compose(r, xt_shfl_down(r, o), k);
    return r;
}

// One CTA per hyper-tile (32 super-tiles = 1024 tiles), one warp per super-tile, one lane per tile: for every
// binade that ALL tiles of the super-tile have a run for, the ordered composition of the 32 runs; then warp 0
// composes the 32 super-tile records the same way into the hyper-tile's record.
__device__ __forceinline__ PaprTileRun xt_pick(const PaprSuperRec &r, int k) // k inside [r.k_lo, r.k_lo + r.nc)
{
    PaprTileRun o = r.r[0];
#pragma unroll
    for (int cd = 1; cd < XT_SUPER_CAND; ++cd)
        if (cd == k - r.k_lo) o = r.r[cd];
    return o;
}

__global__ void __launch_bounds__(1024) papr_xt_compose_kernel(const PaprTileRun *tile_run, const int *tile_code,
                                                               unsigned ntiles, const PaprTileRun *multi_tile,
                                                               PaprSuperRec *super, PaprSuperRec *hyper, u64 *zero_word,
                                                               int ncompose, const PaprFinalizeXArgs fx)
{
    // sharded: one extra CTA exchanges the shards' pass-1 states and derives the levels meanwhile (it spins for the
    // other ranks; the composing CTAs neither wait for it nor are waited for)
    if ((int)blockIdx.x >= ncompose) {
        finalize_levels_x_body(fx);
        return;
    }
    __shared__ PaprSuperRec s_rec[32];
    if (zero_word && blockIdx.x == 0 && threadIdx.x == 0) *zero_word = 0; // the status word of the counts that follow (RES_* bits are OR-ed in)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nsuper = (ntiles + XT_SUPER_TILES - 1) / XT_SUPER_TILES;
    const unsigned nhyper = (nsuper + XT_HYPER_SUPERS - 1) / XT_HYPER_SUPERS;
    for (unsigned hy = blockIdx.x; hy < nhyper; hy += (unsigned)ncompose) {
        const unsigned st = hy * XT_HYPER_SUPERS + warp;
        {
            const unsigned t = st * XT_SUPER_TILES + lane;
            const bool have = st < nsuper && t < ntiles;
            const int code = have ? tile_code[t] : 0;
            PaprTileRun run;
            run.e0 = run.e1 = 0.0;
            if (have) run = tile_run[t];
            const int nc = have ? XT_CODE_NC(code) : 0, K = XT_CODE_K(code);
            const double asum = run.e0;
            // candidate binades of this tile; an all-zero tile (or none at all) is the identity in every binade
            const bool ident = !have || (!(code & XT_CODE_LITERAL) && asum == 0.0);
            int lo = ident ? -5000 : (nc >= 1 ? K : 5000), hi = ident ? 5000 : (nc >= 1 ? K + nc - 1 : -5000);
            lo = __reduce_max_sync(FULL, lo);
            hi = __reduce_min_sync(FULL, hi);
            const int ncs = lo == -5000 ? 0 : max(0, min(hi - lo + 1, XT_SUPER_CAND)); // all identity: nothing to record
            PaprSuperRec s;
            s.k_lo = lo; s.nc = ncs;
#pragma unroll
            for (int cd = 0; cd < XT_SUPER_CAND; ++cd) {
                s.r[cd].e0 = s.r[cd].e1 = 0.0;
                if (cd < ncs) {
                    const int k = lo + cd;
                    PaprTileRun r;
                    r.e0 = r.e1 = 0.0;
                    if (!ident) r = nc == 1 ? run : multi_tile[(size_t)XT_CODE_SLOT(code) * XT_MAX_CAND + (k - K)];
                    s.r[cd] = xt_warp_compose(r, k);
                }
            }
            s.asum = warp_sum_fixed(asum);
            if (lane == 0) {
                if (st < nsuper) super[st] = s;
                s_rec[warp] = s;
            }
        }
        __syncthreads();
        if (warp == 0) { // the same one level up: lane = super-tile
            const PaprSuperRec me = s_rec[lane];
            const bool ident = me.asum == 0.0 && me.nc == 0; // (only zeros, or beyond the end)
            int lo = ident ? -5000 : (me.nc >= 1 ? me.k_lo : 5000), hi = ident ? 5000 : (me.nc >= 1 ? me.k_lo + me.nc - 1 : -5000);
            lo = __reduce_max_sync(FULL, lo);
            hi = __reduce_min_sync(FULL, hi);
            const int ncs = lo == -5000 ? 0 : max(0, min(hi - lo + 1, XT_SUPER_CAND));
            PaprSuperRec h;
            h.k_lo = lo; h.nc = ncs;
#pragma unroll
            for (int cd = 0; cd < XT_SUPER_CAND; ++cd) {
                h.r[cd].e0 = h.r[cd].e1 = 0.0;
                if (cd < ncs) {
                    PaprTileRun r;
                    r.e0 = r.e1 = 0.0;
                    if (!ident) r = xt_pick(me, lo + cd);
                    h.r[cd] = xt_warp_compose(r, lo + cd);
                }
            }
            h.asum = warp_sum_fixed(me.asum);
            if (lane == 0) hyper[hy] = h;
        }
        __syncthreads();
    }
}

void papr_launch_xt_compose(const PaprTileRun *tile_run, const int *tile_code, unsigned ntiles, const PaprTileRun *multi_tile,
                            PaprSuperRec *super, PaprSuperRec *hyper, int grid, cudaStream_t s, unsigned long long *zero_word,
                            const PaprFinalizeXArgs *fx)
{
    PaprFinalizeXArgs none = {};
    papr_xt_compose_kernel<<<grid + (fx ? 1 : 0), 1024, 0, s>>>(tile_run, tile_code, ntiles, multi_tile, super, hyper, zero_word,
                                                                grid, fx ? *fx : none);
}

// ------------------------------------------------------------------------------------------------
// the chain: one CTA of 1024 threads
// ------------------------------------------------------------------------------------------------
enum { // PaprChainList.why
    XW_NONE = 0, XW_TOO_MANY_CROSSINGS = 1, XW_SUPER_MISPREDICTED = 2, XW_TILE_NO_CANDIDATE = 3, XW_TWO_CROSSINGS_IN_TILE = 4,
    XW_BATCH_NOT_FOUND = 5, XW_LANE_NOT_FOUND = 6, XW_TOO_MANY_ITEMS = 7, XW_WALK_BINADE = 8, XW_WALK_OVERFLOW = 9,
    XW_WALK_LITERAL = 10, XW_WALK_ABS = 11, XW_START_UNKNOWN = 12, XW_DECLINED = 13,
};

struct XtCtx {
    const PaprSuperRec *hyper;     // [nhyper] 32 super-tiles each
    unsigned nhyper;
    const PaprSuperRec *super;
    const PaprTileRun *tile_run;
    const int *tile_code;
    const PaprTileRun *multi, *multi_tile;
    unsigned ntiles, nsuper;
    const float *iq;               // the shard (for the 8 + 248 samples around each crossing)
    unsigned long long nsamples;
};

__device__ __forceinline__ double xt_warp_excl_scan(double v, int lane, double *total)
{
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += u;
    }
    *total = __shfl_sync(FULL, inc, 31);
    return inc - v; // exclusive (approximate sums only)
}

// thread 0 of the CTA: apply a shard's items (in shared or global memory) in order.  Returns the status.
__device__ int xt_walk(const PaprChainItem *item, int n, double *state, int *why)
{
    double s = *state;
    for (int i = 0; i < n; ++i) {
        const PaprChainItem it = item[i];
        if (it.type == XT_IT_SEG) {
            if (it.e0 == 0.0 && it.e1 == 0.0) continue; // nothing but zeros: the identity in any binade
            if (xt_expo(s) != it.k) { *why = XW_WALK_BINADE; return XT_FALLBACK; }
            const double inc = xt_lsb(s) ? it.e1 : it.e0;
            const double s2 = __dadd_rn(s, inc); // whole ulps on both sides: exact while it stays in the binade
            if (!(inc >= 0.0) || !(s2 < xt_base(it.k + 1, 0))) { *why = XW_WALK_OVERFLOW; return XT_FALLBACK; }
            s = s2;
        } else if (it.type == XT_IT_LIT) {
            s = __dadd_rn(s, it.e0); // papr.c:104 as written: the add that carries the sum into the next binade
            if (xt_expo(s) != it.k) { *why = XW_WALK_LITERAL; return XT_FALLBACK; }
        } else {
            if (s != 0.0) { *why = XW_WALK_ABS; return XT_FALLBACK; }
            s = it.e0;
        }
    }
    *state = s;
    return XT_OK;
}

#define XT_CHAIN_T 1024
#define XT_MAX_XTILES 64

// this lane's tile inside a crossing super-tile, loaded once
struct XtTile {
    int code;
    PaprTileRun run;             // single tiles: the run; otherwise e0 = approximate sum
    PaprTileRun cand[XT_MAX_CAND]; // multi tiles: the run per candidate
};

__device__ __forceinline__ bool xt_tile_pick(const XtTile &t, bool have, int k, PaprTileRun *out)
{
    out->e0 = out->e1 = 0.0;
    if (!have) return true;
    const int nc = XT_CODE_NC(t.code), K = XT_CODE_K(t.code);
    if (nc == 1 && K == k) { *out = t.run; return true; }
    if (nc > 1 && k >= K && k < K + nc) {
#pragma unroll
        for (int cd = 0; cd < XT_MAX_CAND; ++cd)
            if (cd == k - K) *out = t.cand[cd];
        return true;
    }
    return !(t.code & XT_CODE_LITERAL) && t.run.e0 == 0.0; // an all-zero tile is the identity in any binade
}

// the merged chain of a shard, in shared memory while it is built and walked
struct XtShared {
    PaprChainItem item[XT_MAX_ITEMS];
    int n;
};

// Builds the item list of this shard (sh->item[0..n) in file order, runs of the same binade merged; copied to
// *out as well).  pre_approx = approximate running sum before the shard (0 for the first).  All threads
// return the status.
__device__ int xt_chain_prepare(const XtCtx &c, double pre_approx, PaprChainList *out, XtShared *sh)
{
    __shared__ double s_scan[XT_CHAIN_T];
    __shared__ PaprTileRun s_piece[XT_CHAIN_T + XT_MAX_CROSS + 1];
    __shared__ short s_piece_k[XT_CHAIN_T + XT_MAX_CROSS + 1]; // the binade each piece was composed for
    __shared__ unsigned s_xh[XT_MAX_CROSS];    // crossing hyper-tiles, ascending
    __shared__ double s_xhp[XT_MAX_CROSS];     // approximate running sum at their start
    __shared__ unsigned s_xs[XT_MAX_CROSS];    // crossing super-tiles
    __shared__ double s_xp[XT_MAX_CROSS];
    __shared__ unsigned s_xt[XT_MAX_XTILES];   // crossing tiles
    __shared__ double s_xtp[XT_MAX_XTILES];
    __shared__ int s_xtc[XT_MAX_XTILES];       // ... and their codes
    __shared__ float s_run[XT_CHAIN_T / 32][XT_RUN]; // per warp: the samples of the run in which a crossing happens
    __shared__ unsigned long long s_key[XT_MAX_RAW];
    __shared__ double s_re0[XT_MAX_RAW], s_re1[XT_MAX_RAW]; // the raw items
    __shared__ short s_rk[XT_MAX_RAW];
    __shared__ char s_rtype[XT_MAX_RAW];
    __shared__ short s_order[XT_MAX_RAW];      // raw index of the r-th item in file order
    __shared__ short s_gstart[XT_MAX_ITEMS];
    __shared__ short s_ord2[XT_MAX_RAW];       // ... without the identities
    __shared__ int s_woff[XT_MAX_RAW / 32 + 1];
    __shared__ int s_nxh, s_nx, s_nxt, s_nraw, s_status, s_why, s_ngroups, s_ncomp;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#ifdef XT_TIMING
    unsigned long long tm[12];
    int tmi = 0;
#define XT_MARK() do { if (t == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tm[tmi])); ++tmi; } } while (0)
#else
#define XT_MARK() do { } while (0)
#endif
    XT_MARK();
    if (t == 0) { s_nxh = 0; s_nx = 0; s_nxt = 0; s_nraw = 0; s_status = XT_OK; s_why = XW_NONE; s_ngroups = 0; s_ncomp = 0; }
    __syncthreads();
    auto fail = [&](int why) { if (atomicCAS(&s_status, XT_OK, XT_FALLBACK) == XT_OK) s_why = why; };
    auto emit = [&](int type, int k, unsigned long long pos, double e0, double e1) -> int {
        const int i = atomicAdd(&s_nraw, 1);
        if (i >= XT_MAX_RAW) { fail(XW_TOO_MANY_ITEMS); return -1; }
        s_key[i] = pos; s_rtype[i] = (char)type; s_rk[i] = (short)k; s_re0[i] = e0; s_re1[i] = e1;
        return i;
    };

    // ---- 1. approximate running sum at every hyper-tile (block scan of chunk sums)
    const unsigned m = (c.nhyper + XT_CHAIN_T - 1) / XT_CHAIN_T;
    const unsigned lo = min(c.nhyper, (unsigned)t * m), hi = min(c.nhyper, lo + m);
    double chunk = 0.0;
    for (unsigned i = lo; i < hi; ++i) chunk += c.hyper[i].asum;
    s_scan[t] = chunk;
    __syncthreads();
    for (int o = 1; o < XT_CHAIN_T; o <<= 1) {
        const double u = t >= o ? s_scan[t - o] : 0.0;
        __syncthreads();
        s_scan[t] += u;
        __syncthreads();
    }
    const double total = pre_approx + s_scan[XT_CHAIN_T - 1];
    if (t == 0) out->approx = s_scan[XT_CHAIN_T - 1];
    if (!isfinite(total)) { // NaN / Inf among the powers: so is the reference's sum, nothing to emulate
        if (t == 0) { out->n = 0; out->status = XT_NONFINITE; out->why = XW_NONE; sh->n = 0; }
        __syncthreads();
        return XT_NONFINITE;
    }
    const double p_chunk = pre_approx + (s_scan[t] - chunk);
    XT_MARK(); // 1: prefix done

    // ---- 2. hyper-tiles inside which the running sum changes binade (or starts from zero)
    {
        double p = p_chunk;
        for (unsigned i = lo; i < hi; ++i) {
            const double pn = p + c.hyper[i].asum;
            if (xt_expo(p) != xt_expo(pn)) {
                const int j = atomicAdd(&s_nxh, 1);
                if (j < XT_MAX_CROSS) { s_xh[j] = i; s_xhp[j] = p; } else fail(XW_TOO_MANY_CROSSINGS);
            }
            p = pn;
        }
    }
    __syncthreads();
    const int nxh = min(s_nxh, XT_MAX_CROSS);
    if (t == 0) // ascending (<= 44 entries)
        for (int i = 1; i < nxh; ++i) {
            const unsigned v = s_xh[i];
            const double pv = s_xhp[i];
            int j = i - 1;
            while (j >= 0 && s_xh[j] > v) { s_xh[j + 1] = s_xh[j]; s_xhp[j + 1] = s_xhp[j]; --j; }
            s_xh[j + 1] = v; s_xhp[j + 1] = pv;
        }
    __syncthreads();

    // ---- 3. the stretches of hyper-tiles between them: every thread composes its chunk, cut at the crossing
    //         hyper-tiles; piece index t + (crossings before it) puts all pieces in file order, and the pieces
    //         of stretch number s are exactly those with index - thread == s
    {
        int xb = 0; // crossing hyper-tiles before this chunk
        while (xb < nxh && s_xh[xb] < lo) ++xb;
        int seg = xb;
        double p = p_chunk;
        PaprTileRun cur;
        cur.e0 = cur.e1 = 0.0;
        int k = xt_expo(p);
        for (unsigned i = lo; i < hi; ++i) {
            const PaprSuperRec sr = c.hyper[i];
            if (seg < nxh && s_xh[seg] == i) {
                s_piece[t + seg] = cur;
                s_piece_k[t + seg] = (short)k;
                cur.e0 = cur.e1 = 0.0;
                ++seg;
                p += sr.asum;
                k = xt_expo(p);
                continue;
            }
            if (sr.nc > 0 && k >= sr.k_lo && k < sr.k_lo + sr.nc) cur = xt_compose(cur, xt_pick(sr, k), k);
            else if (!(sr.asum == 0.0)) fail(XW_SUPER_MISPREDICTED); // (nothing but zeros: the identity in any binade)
            p += sr.asum;
        }
        s_piece[t + seg] = cur;
        s_piece_k[t + seg] = (short)k;
    }
    __syncthreads();
    for (int s = warp; s <= nxh; s += XT_CHAIN_T / 32) { // one warp per stretch
        const unsigned a = s == 0 ? 0u : s_xh[s - 1] + 1u, b = s == nxh ? c.nhyper : s_xh[s];
        if (a >= b) continue;
        // binade of the stretch: the running sum at its start
        const double p0 = s == 0 ? pre_approx : s_xhp[s - 1] + c.hyper[s_xh[s - 1]].asum;
        const int k = xt_expo(p0);
        const unsigned ta = a / m, tb = (b - 1) / m;
        const unsigned cnt = tb - ta + 1, per = (cnt + 31) / 32; // lanes take contiguous thread ranges
        PaprTileRun r;
        r.e0 = r.e1 = 0.0;
        for (unsigned q = 0; q < per; ++q) {
            const unsigned th = ta + lane * per + q;
            if (th <= tb) {
                const PaprTileRun pc = s_piece[th + s];
                // a piece composed for another binade (the approximate running sum read differently by two
                // threads right at a power of two) must not be mixed in
                if (s_piece_k[th + s] != (short)k && !(pc.e0 == 0.0 && pc.e1 == 0.0)) fail(XW_SUPER_MISPREDICTED);
                r = xt_compose(r, pc, k);
            }
        }
        r = xt_warp_compose(r, k);
        if (lane == 0) emit(XT_IT_SEG, k, (unsigned long long)a * XT_HYPER_SAMPLES, r.e0, r.e1);
    }

    __syncthreads();
    XT_MARK(); // 2: stretches of hyper-tiles done
    // ---- 3b. inside every crossing hyper-tile: the super-tiles in which it happens, the stretches between
    for (int x = warp; x < nxh; x += XT_CHAIN_T / 32) {
        const unsigned hy = s_xh[x], st = hy * XT_HYPER_SUPERS + lane;
        const bool have = st < c.nsuper;
        PaprSuperRec me;
        me.asum = 0.0; me.k_lo = 0; me.nc = 0;
#pragma unroll
        for (int cd = 0; cd < XT_SUPER_CAND; ++cd) me.r[cd].e0 = me.r[cd].e1 = 0.0;
        if (have) me = c.super[st];
        double tot;
        const double pl = s_xhp[x] + xt_warp_excl_scan(me.asum, lane, &tot), pa = pl + me.asum;
        const bool cross = have && xt_expo(pl) != xt_expo(pa);
        if (cross) {
            const int j = atomicAdd(&s_nx, 1);
            if (j < XT_MAX_CROSS) { s_xs[j] = st; s_xp[j] = pl; } else fail(XW_TOO_MANY_CROSSINGS);
        }
        const unsigned bnd = __ballot_sync(FULL, cross);
        int a = 0;
        while (a < 32) {
            const unsigned rest = bnd >> a;
            const int b = rest ? a + (__ffs(rest) - 1) : 32; // stretch = lanes [a, b)
            if (b > a) {
                const int k = xt_expo(__shfl_sync(FULL, pl, a)); // binade at its start
                PaprTileRun r;
                r.e0 = r.e1 = 0.0;
                bool ok = true;
                if (have && lane >= a && lane < b) {
                    if (me.nc > 0 && k >= me.k_lo && k < me.k_lo + me.nc) r = xt_pick(me, k);
                    else ok = me.asum == 0.0;
                }
                if (!__all_sync(FULL, ok)) { if (lane == 0) fail(XW_SUPER_MISPREDICTED); }
                r = xt_warp_compose(r, k);
                if (lane == 0) emit(XT_IT_SEG, k, (unsigned long long)(hy * XT_HYPER_SUPERS + a) * XT_SUPER_SAMPLES, r.e0, r.e1);
            }
            a = b + 1;
        }
    }
    __syncthreads();
    const int nx = min(s_nx, XT_MAX_CROSS);
    XT_MARK(); // 3: 3b done

    // ---- 4. inside every crossing super-tile: the tiles in which it happens, and the stretches of tiles between
    for (int x = warp; x < nx; x += XT_CHAIN_T / 32) {
        const unsigned st = s_xs[x], tl = st * XT_SUPER_TILES + lane;
        const bool have = tl < c.ntiles;
        XtTile tt;
        tt.code = have ? c.tile_code[tl] : 0;
        tt.run.e0 = tt.run.e1 = 0.0;
        if (have) tt.run = c.tile_run[tl];
#pragma unroll
        for (int cd = 0; cd < XT_MAX_CAND; ++cd) {
            tt.cand[cd].e0 = tt.cand[cd].e1 = 0.0;
            if (have && cd < XT_CODE_NC(tt.code) && XT_CODE_NC(tt.code) > 1)
                tt.cand[cd] = c.multi_tile[(size_t)XT_CODE_SLOT(tt.code) * XT_MAX_CAND + cd];
        }
        const double asum = tt.run.e0;
        double tot;
        const double pl = s_xp[x] + xt_warp_excl_scan(asum, lane, &tot), pa = pl + asum;
        const bool lit = have && (tt.code & XT_CODE_LITERAL);
        const bool cross = have && !lit && xt_expo(pl) != xt_expo(pa);
        if (cross && xt_expo(pa) != xt_expo(pl) + 1) fail(xt_expo(pl) <= -4000 ? XW_START_UNKNOWN : XW_TWO_CROSSINGS_IN_TILE);
        if (lit) {
            if (pl != 0.0) fail(XW_WALK_ABS);
            emit(XT_IT_ABS, 0, (unsigned long long)tl * XT_TILE_SAMPLES, tt.run.e0, 0.0);
        }
        if (cross) {
            const int j = atomicAdd(&s_nxt, 1);
            if (j < XT_MAX_XTILES) { s_xt[j] = tl; s_xtp[j] = pl; s_xtc[j] = tt.code; } else fail(XW_TOO_MANY_CROSSINGS);
        }
        // the samples of a crossing tile will be needed in step 5 (they left L2 long ago): start fetching them now
        {
            unsigned xm = __ballot_sync(FULL, cross);
            while (xm) {
                const int q = __ffs(xm) - 1;
                xm &= xm - 1;
                const unsigned long long s0 = (unsigned long long)(st * XT_SUPER_TILES + q) * XT_TILE_SAMPLES;
                const int qc = __shfl_sync(FULL, tt.code, q);
                if (XT_CODE_NC(qc) > 1 && lane < XT_MAX_CAND) // ... and its batch records (128 B per candidate)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(c.multi + ((size_t)XT_CODE_SLOT(qc) * XT_MAX_CAND + lane) * XT_TILE_BATCHES));
                const char *base = reinterpret_cast<const char *>(c.iq + 2 * s0) + 1024 * lane; // 32 KiB per tile: 1 KiB per lane
#pragma unroll
                for (int l = 0; l < 8; ++l)
                    if (s0 + 128ull * lane + 16ull * l < c.nsamples)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + 128 * l));
            }
        }
        // stretches of ordinary tiles between the boundaries (crossing / literal tiles)
        const unsigned bnd = __ballot_sync(FULL, cross || lit);
        int a = 0;
        while (a < 32) {
            const unsigned rest = bnd >> a;
            const int b = rest ? a + (__ffs(rest) - 1) : 32; // stretch = lanes [a, b)
            if (b > a) {
                const int k = xt_expo(__shfl_sync(FULL, pl, a)); // binade at its start
                PaprTileRun r;
                const bool ok = xt_tile_pick(tt, have && lane >= a && lane < b, k, &r);
                if (!__all_sync(FULL, ok)) { if (lane == 0) fail(XW_TILE_NO_CANDIDATE); }
                r = xt_warp_compose(r, k);
                if (lane == 0) emit(XT_IT_SEG, k, (unsigned long long)(st * XT_SUPER_TILES + a) * XT_TILE_SAMPLES, r.e0, r.e1);
            }
            a = b + 1;
        }
    }
    __syncthreads();

    XT_MARK(); // 4: tiles done
    // ---- 5. inside every crossing tile: the batch, then the XT_RUN samples
    const int nxt = min(s_nxt, XT_MAX_XTILES);
    for (int x = warp; x < nxt; x += XT_CHAIN_T / 32) {
        const unsigned tl = s_xt[x];
        const int code = s_xtc[x], nc = XT_CODE_NC(code), K = XT_CODE_K(code);
        const int kb = xt_expo(s_xtp[x]), ka = kb + 1;
        if (nc < 2 || kb < K || ka >= K + nc) { if (lane == 0) fail(XW_TILE_NO_CANDIDATE); continue; }
        const PaprTileRun *mb = c.multi + ((size_t)XT_CODE_SLOT(code) * XT_MAX_CAND + (kb - K)) * XT_TILE_BATCHES;
        const PaprTileRun *ma = mb + XT_TILE_BATCHES;
        PaprTileRun rb, ra;
        rb.e0 = rb.e1 = ra.e0 = ra.e1 = 0.0;
        if (lane < XT_TILE_BATCHES) { rb = mb[lane]; ra = ma[lane]; }
        double tot;
        const double pl = s_xtp[x] + xt_warp_excl_scan(rb.e0, lane, &tot);
        const unsigned xb = __ballot_sync(FULL, lane < XT_TILE_BATCHES && xt_expo(pl) != xt_expo(pl + rb.e0));
        if (__popc(xb) != 1) { if (lane == 0) fail(XW_BATCH_NOT_FOUND); continue; }
        const int bx = __ffs(xb) - 1;
        const double pbx = __shfl_sync(FULL, pl, bx);
        const unsigned long long tile_pos = (unsigned long long)tl * XT_TILE_SAMPLES;
        // the crossing batch: this lane's XT_RUN samples (issued before the compositions below: HBM latency)
        const unsigned long long s0 = tile_pos + (unsigned long long)bx * XT_BATCH_SAMPLES + (unsigned long long)XT_RUN * lane;
        float v[XT_RUN];
#pragma unroll
        for (int j = 0; j < XT_RUN; ++j) {
            v[j] = 0.f;
            if (s0 + j < c.nsamples) {
                const float2 h = *reinterpret_cast<const float2 *>(c.iq + 2 * (s0 + j));
                v[j] = power_of(h.x, h.y);
            }
        }
        { // batches before: binade kb; after: ka
            PaprTileRun r = rb;
            if (lane >= bx) r.e0 = r.e1 = 0.0;
            r = xt_warp_compose(r, kb);
            PaprTileRun q = ra;
            if (lane <= bx || lane >= XT_TILE_BATCHES) q.e0 = q.e1 = 0.0;
            q = xt_warp_compose(q, ka);
            if (lane == 0) {
                emit(XT_IT_SEG, kb, tile_pos + 1, r.e0, r.e1); // after the tile-level stretch that ends here
                emit(XT_IT_SEG, ka, tile_pos + (unsigned long long)(bx + 1) * XT_BATCH_SAMPLES, q.e0, q.e1);
            }
        }
        PaprTileRun lb, la;
        {
            double a0 = xt_base(kb, 0), a1 = xt_base(kb, 1), b0 = xt_base(ka, 0), b1 = xt_base(ka, 1);
#pragma unroll
            for (int j = 0; j < XT_RUN; ++j) {
                const double d = (double)v[j];
                a0 = __dadd_rn(a0, d); a1 = __dadd_rn(a1, d);
                b0 = __dadd_rn(b0, d); b1 = __dadd_rn(b1, d);
            }
            lb.e0 = __dsub_rn(a0, xt_base(kb, 0)); lb.e1 = __dsub_rn(a1, xt_base(kb, 1));
            la.e0 = __dsub_rn(b0, xt_base(ka, 0)); la.e1 = __dsub_rn(b1, xt_base(ka, 1));
        }
        const double ql = pbx + xt_warp_excl_scan(lb.e0, lane, &tot);
        const unsigned xl = __ballot_sync(FULL, xt_expo(ql) != xt_expo(ql + lb.e0));
        if (__popc(xl) != 1) { if (lane == 0) fail(XW_LANE_NOT_FOUND); continue; }
        const int lx = __ffs(xl) - 1;
        {
            PaprTileRun r = lb;
            if (lane >= lx) r.e0 = r.e1 = 0.0;
            r = xt_warp_compose(r, kb);
            PaprTileRun q = la;
            if (lane <= lx) q.e0 = q.e1 = 0.0;
            q = xt_warp_compose(q, ka);
            const unsigned long long bpos = tile_pos + (unsigned long long)bx * XT_BATCH_SAMPLES;
            if (lane == 0) {
                emit(XT_IT_SEG, kb, bpos + 2, r.e0, r.e1);
                emit(XT_IT_SEG, ka, bpos + (unsigned long long)XT_RUN * (lx + 1), q.e0, q.e1);
            }
            // inside the run: the one sample whose add leaves binade kb is applied literally; the samples before
            // it form a run in kb, those after it a run in ka.  Lane j takes sample j of the run.
            if (lane == lx)
#pragma unroll
                for (int j = 0; j < XT_RUN; ++j) s_run[warp][j] = v[j];
            __syncwarp();
            const double d = lane < XT_RUN ? (double)s_run[warp][lane] : 0.0;
            const double qx = __shfl_sync(FULL, ql, lx);
            const double qs = qx + xt_warp_excl_scan(d, lane, &tot);
            const unsigned xj = __ballot_sync(FULL, lane < XT_RUN && xt_expo(qs) != xt_expo(qs + d));
            if (__popc(xj) != 1) { if (lane == 0) fail(XW_LANE_NOT_FOUND); continue; }
            const int jx = __ffs(xj) - 1;
            PaprTileRun pre, post;
            pre.e0 = pre.e1 = post.e0 = post.e1 = 0.0;
            if (lane < jx) {
                pre.e0 = __dsub_rn(__dadd_rn(xt_base(kb, 0), d), xt_base(kb, 0));
                pre.e1 = __dsub_rn(__dadd_rn(xt_base(kb, 1), d), xt_base(kb, 1));
            } else if (lane > jx && lane < XT_RUN) {
                post.e0 = __dsub_rn(__dadd_rn(xt_base(ka, 0), d), xt_base(ka, 0));
                post.e1 = __dsub_rn(__dadd_rn(xt_base(ka, 1), d), xt_base(ka, 1));
            }
            pre = xt_warp_compose(pre, kb);
            post = xt_warp_compose(post, ka);
            const double lit = __shfl_sync(FULL, d, jx);
            if (lane == 0) {
                const unsigned long long rpos = bpos + (unsigned long long)XT_RUN * lx;
                emit(XT_IT_SEG, kb, rpos + 3, pre.e0, pre.e1);
                emit(XT_IT_LIT, ka, rpos + 4, lit, 0.0);
                emit(XT_IT_SEG, ka, rpos + 5, post.e0, post.e1);
            }
            __syncwarp();
        }
    }
    __threadfence_block();
    __syncthreads();

    XT_MARK(); // 5: batches + lanes done
    // ---- 6. file order: rank every item by its key
    const int nr = min(s_nraw, XT_MAX_RAW);
    for (int i = t; i < nr; i += XT_CHAIN_T) {
        const unsigned long long ki = s_key[i];
        int rank = 0;
        for (int j = 0; j < nr; ++j) rank += (s_key[j] < ki) || (s_key[j] == ki && j < i);
        s_order[rank] = (short)i;
    }
    __syncthreads();
    // ---- 7. identities dropped; runs of the same binade that follow each other become one item
    {
        // (a) compaction: position r (file order) -> position among the items that do something
        const int i = t < nr ? s_order[t] : 0;
        const bool keep = t < nr && !(s_rtype[i] == XT_IT_SEG && s_re0[i] == 0.0 && s_re1[i] == 0.0);
        const unsigned kb = __ballot_sync(FULL, keep);
        if (lane == 0 && warp <= XT_MAX_RAW / 32) s_woff[warp] = __popc(kb);
        __syncthreads();
        if (t == 0) {
            int acc = 0;
            for (int w = 0; w <= XT_MAX_RAW / 32; ++w) { const int c2 = s_woff[w]; s_woff[w] = acc; acc += c2; }
            s_ncomp = acc;
        }
        __syncthreads();
        if (keep) s_ord2[s_woff[warp] + __popc(kb & ((1u << lane) - 1u))] = (short)i;
        __syncthreads();
        // (b) group heads: an item starts a group unless it and its predecessor are runs of the same binade
        const int nc2 = s_ncomp;
        const int me = t < nc2 ? s_ord2[t] : 0, prev = (t > 0 && t < nc2) ? s_ord2[t - 1] : 0;
        const bool head = t < nc2 && !(t > 0 && s_rtype[me] == XT_IT_SEG && s_rtype[prev] == XT_IT_SEG && s_rk[me] == s_rk[prev]);
        const unsigned hb = __ballot_sync(FULL, head);
        if (lane == 0 && warp <= XT_MAX_RAW / 32) s_woff[warp] = __popc(hb);
        __syncthreads();
        if (t == 0) {
            int acc = 0;
            for (int w = 0; w <= XT_MAX_RAW / 32; ++w) { const int c2 = s_woff[w]; s_woff[w] = acc; acc += c2; }
            s_ngroups = acc;
            if (acc > XT_MAX_ITEMS) fail(XW_TOO_MANY_ITEMS);
        }
        __syncthreads();
        if (head) {
            const int g = s_woff[warp] + __popc(hb & ((1u << lane) - 1u));
            if (g < XT_MAX_ITEMS) s_gstart[g] = (short)t;
        }
        __syncthreads();
    }
    XT_MARK(); // 6: sorted + grouped
    const int ng = min(s_ngroups, XT_MAX_ITEMS);
    for (int g = t; g < ng; g += XT_CHAIN_T) {
        const int q0 = s_gstart[g], q1 = g + 1 < ng ? s_gstart[g + 1] : s_ncomp;
        const int i0 = s_ord2[q0];
        PaprChainItem it;
        it.type = s_rtype[i0]; it.k = s_rk[i0];
        it.e0 = s_re0[i0]; it.e1 = s_re1[i0];
        if (it.type == XT_IT_SEG) {
            PaprTileRun r;
            r.e0 = it.e0; r.e1 = it.e1;
            for (int q = q0 + 1; q < q1; ++q) {
                const int iq = s_ord2[q];
                PaprTileRun b2;
                b2.e0 = s_re0[iq]; b2.e1 = s_re1[iq];
                r = xt_compose(r, b2, it.k);
            }
            it.e0 = r.e0; it.e1 = r.e1;
        }
        sh->item[g] = it;
        out->item[g] = it; // a copy in global memory (other ranks walk it too in a sharded analysis)
    }
    if (t == 0) { sh->n = ng; out->n = ng; out->status = s_status; out->why = s_why; }
    __threadfence();
    __syncthreads();
    XT_MARK(); // 7: merged + copied
#ifdef XT_TIMING
    if (t == 0) {
        printf("xt chain ns: prefix %llu stretches %llu hyper %llu tiles %llu lanes %llu sort %llu merge %llu | raw %d items %d nxh %d nx %d nxt %d\n",
               tm[1] - tm[0], tm[2] - tm[1], tm[3] - tm[2], tm[4] - tm[3], tm[5] - tm[4], tm[6] - tm[5], tm[7] - tm[6], nr, ng, nxh, nx, nxt);
    }
#endif
    return s_status;
}

// papr.c:147-151 restated (as in papr_count_kernel), by all XT_CHAIN_T threads of a CTA: levels j0, j0 + stride, ...
// below L: counts[j] = samples above the cell range + samples in the cells above cell(T_j) + the fine counters of
// the values of that cell that are > T_j.  A threshold outside the planned cells / windows => RES_MISS.
template <class LevelOf>
static __device__ void xt_count_levels(const PaprEpilogueArgs &a, int L, int j0, int stride, LevelOf level_of)
{
    __shared__ u64 s_red[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const PaprPlan pl = *a.plan;
    if (pl.status & PLAN_BSEARCH) return;
    const unsigned fmask = (1u << pl.sh) - 1u;
    for (int j = j0; j < L; j += stride) {
        const unsigned tbits = __float_as_uint(level_of(j));
        const int d = (int)(tbits >> pl.sh) - pl.cell_base;
        const bool in = (pl.status & PLAN_HIST) && (unsigned)d < (unsigned)pl.ncells;
        const unsigned fb = in ? a.fine_base[d] : 0;
        if (!fb) {
            if (t == 0) { atomicOr(a.status_word, (u64)RES_MISS); a.counts[j] = 0; }
            continue;
        }
        const u64 *f = a.g_fine + (fb - 1u);
        u64 acc = 0;
        for (unsigned k = (tbits & fmask) + 1 + t; k <= fmask; k += XT_CHAIN_T) acc += f[k];
        for (int cc = d + 1 + t; cc < pl.ncells; cc += XT_CHAIN_T) acc += a.g_hist[cc];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (warp == 0) {
            acc = s_red[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
            if (lane == 0) a.counts[j] = acc + *a.g_over;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// single shard: the chain SIDE BY SIDE with the scalar epilogue and the level counts (one launch instead of
// chain -> finalize+levels -> count).  CTA 0 chains the sequential sum.  Every other CTA folds the CTA partials
// itself (148 x 72 B), evaluates papr.c:131-141 with the FIXED-ORDER sum of the partials - which differs from
// the sequential sum by ~1e-12 relative, almost never enough to move a float32 level - and counts its share of
// the levels; CTA 1 also stores the statistics and the level table.  The host then derives the levels from the
// chain's exact sum with its own libm, as it always does, and accepts the counts only if the device's levels are
// bit-identical to those; otherwise (about once in 10^4 analyses) the counts are taken again with the right
// levels (papr_engine.cu: recount).  `bias` scales the sum the speculative levels are derived from (test hook).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XT_CHAIN_T) papr_xt_epilogue_kernel(XtCtx c, PaprChainList *out, const PaprEpilogueArgs a)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (blockIdx.x == 0) {
        __shared__ XtShared sh;
        const int st = xt_chain_prepare(c, 0.0, out, &sh);
        if (t == 0) {
            int status = st, why = out->why;
            double s = 0.0;
            if (st == XT_OK) {
                why = XW_NONE;
                status = xt_walk(sh.item, sh.n, &s, &why);
                out->status = status; out->why = why; out->exact = s;
            }
            a.chain_report[0] = status;
            a.chain_report[1] = why;
            *a.chain_exact = s;
        }
        return;
    }
    __shared__ double s_sum[32];
    __shared__ int s_val[PAPR_NTRACK][32];
    __shared__ u64 s_idx[PAPR_NTRACK][32];
    __shared__ double s_avg, s_ratio;
    // ---- fold the CTA partials: fixed order, (value desc, index asc) = first occurrence (papr.c:105-126)
    double sum = 0.0;
    int val[PAPR_NTRACK];
    u64 idx[PAPR_NTRACK];
#pragma unroll
    for (int k = 0; k < PAPR_NTRACK; ++k) { val[k] = 0; idx[k] = 0; }
    for (int i = t; i < a.nctas; i += XT_CHAIN_T) {
        sum += a.wp[i].sum;
#pragma unroll
        for (int k = 0; k < PAPR_NTRACK; ++k) {
            const int v = a.wp[i].val[k];
            const u64 ix = a.wp[i].idx[k];
            if (v > 0 && better(v, ix, val[k], idx[k])) { val[k] = v; idx[k] = ix; }
        }
    }
    for (int round = 0; round < 2; ++round) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_down_sync(FULL, sum, o);
#pragma unroll
            for (int k = 0; k < PAPR_NTRACK; ++k) {
                const int v = __shfl_down_sync(FULL, val[k], o);
                const u64 ix = __shfl_down_sync(FULL, idx[k], o);
                if (v > 0 && better(v, ix, val[k], idx[k])) { val[k] = v; idx[k] = ix; }
            }
        }
        if (round == 0) {
            if (lane == 0) {
                s_sum[warp] = sum;
#pragma unroll
                for (int k = 0; k < PAPR_NTRACK; ++k) { s_val[k][warp] = val[k]; s_idx[k][warp] = idx[k]; }
            }
            __syncthreads();
            sum = s_sum[lane]; // (every warp folds the 32 warp results again: identical everywhere)
#pragma unroll
            for (int k = 0; k < PAPR_NTRACK; ++k) { val[k] = s_val[k][lane]; idx[k] = s_idx[k][lane]; }
        }
    }
    if (t == 0) {
        PaprDevStats m;
        m.sum = sum;
        m.n = a.n;
#pragma unroll
        for (int k = 0; k < PAPR_NTRACK; ++k) { m.val[k] = val[k]; m.idx[k] = idx[k]; }
        m.flags = isfinite(sum) ? 0u : PAPR_FLAG_NONFINITE;
        const double avg = __ddiv_rn(__dmul_rn(sum, a.bias), (double)(long long)a.n);
        s_avg = avg;
        s_ratio = __ddiv_rn((double)__int_as_float(val[TR_PEAK]), avg);
        if (blockIdx.x == 1) {
            *a.local = m;
            *a.merged = m;
            a.lv->avg = avg;
            a.lv->ratio = s_ratio;
            a.lv->graph = a.graph;
        }
    }
    __syncthreads();
    const double avg = s_avg, ratio = s_ratio;
    int L = 0; // number of j with ratio >= ratio_min[j] (non-decreasing table; a NaN ratio reaches no level)
    for (int j0 = 0; j0 < a.tb.nlevels_max; j0 += XT_CHAIN_T)
        L += __syncthreads_count(j0 + t < a.tb.nlevels_max && ratio >= a.tb.ratio_min[j0 + t]);
    if (blockIdx.x == 1) {
        if (t == 0) a.lv->L = L;
        for (int j = t; j < L; j += XT_CHAIN_T) a.lv->level[j] = __double2float_rn(__dmul_rn(a.tb.pow10[j], avg));
    }
    // ---- this CTA's levels
    xt_count_levels(a, L, (int)blockIdx.x - 1, (int)gridDim.x - 1,
                    [&](int j) { return __double2float_rn(__dmul_rn(a.tb.pow10[j], avg)); });
}

void papr_launch_xt_epilogue(const PaprSuperRec *hyper, const PaprSuperRec *super, const PaprTileRun *tile_run, const int *tile_code,
                             const PaprTileRun *multi, const PaprTileRun *multi_tile, unsigned ntiles, const float *iq,
                             unsigned long long nsamples, PaprChainList *out, const PaprEpilogueArgs &a, int grid, cudaStream_t s)
{
    XtCtx c;
    c.hyper = hyper; c.super = super; c.tile_run = tile_run; c.tile_code = tile_code; c.multi = multi; c.multi_tile = multi_tile;
    c.ntiles = ntiles; c.nsuper = (ntiles + XT_SUPER_TILES - 1) / XT_SUPER_TILES; c.iq = iq; c.nsamples = nsamples;
    c.nhyper = (c.nsuper + XT_HYPER_SUPERS - 1) / XT_HYPER_SUPERS;
    papr_xt_epilogue_kernel<<<grid < 2 ? 2 : grid, XT_CHAIN_T, 0, s>>>(c, out, a);
}

// sharded (one process per GPU): every rank builds the list of its own shard - placing it after the lower
// ranks' approximate sums, which the statistics exchange of this analysis has just delivered - publishes it to
// every peer's window over NVLink, collects the others', and walks ALL lists in rank order from a running sum
// of 0: the same work and the same verdict on every rank, no hand-over from GPU to GPU.
// With the level counts side by side, as in the single-shard epilogue: CTA 0 is the chain; the other CTAs count
// this shard's samples above the levels that the statistics exchange (the extra CTA of papr_xt_compose_kernel, or
// papr_finalize_levels_x_kernel) has just derived from the merged FIXED-ORDER sums (a.lv) - the host verifies those
// levels against the ones of the chained sum afterwards; the counting CTA that finishes last exchanges the counts.
// decline != 0: this rank's sweep produced no runs (shard too small for the TMA-fed sweep, or the exact sum is switched
// off): it publishes an empty list marked XT_FALLBACK, so that EVERY rank reports the fall-back - no agreement between
// the ranks is needed beforehand, and the exchange sequence is the same whatever the shard sizes are.
__global__ void __launch_bounds__(XT_CHAIN_T) papr_xt_epilogue_x_kernel(XtCtx c, PaprChainList *out, PaprPlan *plan,
                                                                       PaprPeers pp, u64 seq, const PaprEpilogueArgs a,
                                                                       int decline, u64 seq_counts)
{
    if (blockIdx.x != 0) {
        const PaprDevLevels *lv = a.lv;
        const int L = min(max(lv->L, 0), PAPR_MAX_LEVELS);
        xt_count_levels(a, L, (int)blockIdx.x - 1, (int)gridDim.x - 1, [&](int j) { return lv->level[j]; });
        // the counting CTA that finishes last exchanges the counts with the other ranks (no separate launch)
        __shared__ unsigned s_ticket;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = atomicAdd(a.done, 1u);
        __syncthreads();
        if (s_ticket == gridDim.x - 2) {
            if (threadIdx.x == 0) *a.done = 0; // ready for the next launch
            xchg_counts(a.counts, L, plan, pp, seq_counts);
        }
        return;
    }
    __shared__ XtShared sh;
    __shared__ int s_st, s_why;
    __shared__ double s_state;
    double pre = 0.0;
    for (int q = 0; q < pp.rank; ++q)
        pre += __longlong_as_double((long long)ld_volatile(reinterpret_cast<const u64 *>(&pp.win[pp.rank]->slot[q].stats.sum)));
    if (decline) {
        if (threadIdx.x == 0) { out->n = 0; out->status = XT_FALLBACK; out->why = XW_DECLINED; out->approx = 0.0; sh.n = 0; }
        __threadfence();
        __syncthreads();
    } else {
        xt_chain_prepare(c, pre, out, &sh);
    }
    // header + the items in use
    const int words = (int)((offsetof(PaprChainList, item) + sizeof(PaprChainItem) * (size_t)max(out->n, 0)) / 8);
    xchg_publish(pp, XK_CHAIN, offsetof(PaprXchgSlot, chain), reinterpret_cast<const u64 *>(out), words, seq);
    const bool ok = xchg_wait(pp, XK_CHAIN, seq);
    if (threadIdx.x == 0) {
        if (!ok) plan->pad = 1;
        s_st = ok ? XT_OK : XT_FALLBACK;
        s_why = XW_NONE;
        s_state = 0.0;
    }
    __syncthreads();
    for (int q = 0; q < pp.world; ++q) {
        const PaprChainList *src = &pp.win[pp.rank]->slot[q].chain;
        const int nq = min((int)ld_volatile(reinterpret_cast<const u64 *>(src)) & 0x7fffffff, XT_MAX_ITEMS); // {n, status}
        const int stq = (int)(ld_volatile(reinterpret_cast<const u64 *>(src)) >> 32);
        for (int i = threadIdx.x; i < nq * (int)(sizeof(PaprChainItem) / 8); i += XT_CHAIN_T)
            reinterpret_cast<u64 *>(sh.item)[i] = ld_volatile(reinterpret_cast<const u64 *>(src->item) + i);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (stq == XT_NONFINITE) { if (s_st == XT_OK) s_st = XT_NONFINITE; }
            else if (stq != XT_OK) { if (s_st != XT_FALLBACK) { s_st = XT_FALLBACK; s_why = (int)ld_volatile(reinterpret_cast<const u64 *>(src) + 1) & 0xffff; } }
            else if (s_st == XT_OK) {
                double s = s_state;
                int why = XW_NONE;
                s_st = xt_walk(sh.item, nq, &s, &why);
                s_why = why;
                s_state = s;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out->status = s_st; out->why = s_why; out->exact = s_state;
        a.chain_report[0] = s_st;
        a.chain_report[1] = s_why;
        *a.chain_exact = s_state;
    }
}

void papr_launch_xt_epilogue_x(const PaprSuperRec *hyper, const PaprSuperRec *super, const PaprTileRun *tile_run,
                               const int *tile_code, const PaprTileRun *multi, const PaprTileRun *multi_tile, unsigned ntiles,
                               const float *iq, unsigned long long nsamples, PaprChainList *out, PaprPlan *plan, PaprPeers pp,
                               unsigned long long seq, const PaprEpilogueArgs &a, int grid, cudaStream_t s, int decline,
                               unsigned long long seq_counts)
{
    XtCtx c;
    c.hyper = hyper; c.super = super; c.tile_run = tile_run; c.tile_code = tile_code; c.multi = multi; c.multi_tile = multi_tile;
    c.ntiles = ntiles; c.nsuper = (ntiles + XT_SUPER_TILES - 1) / XT_SUPER_TILES; c.iq = iq; c.nsamples = nsamples;
    c.nhyper = (c.nsuper + XT_HYPER_SUPERS - 1) / XT_HYPER_SUPERS;
    papr_xt_epilogue_x_kernel<<<grid < 2 ? 2 : grid, XT_CHAIN_T, 0, s>>>(c, out, plan, pp, seq, a, decline, seq_counts);
}
