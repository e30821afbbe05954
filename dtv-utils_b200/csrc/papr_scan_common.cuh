// papr_scan_common.cuh — device helpers shared by the scan kernels (papr_kernels.cu: the LDG-fed
// statistics / CCDF kernels; papr_exact.cu: the TMA-fed fused kernel that also emulates the
// reference's sequential double sum).  Per-sample work of drmpeg/dtv-utils papr.c:103-126,147-151.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "papr_device.cuh"

#define FULL 0xffffffffu
typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
// 16-byte streaming load: read-only path, do not allocate in L1 (every byte is used exactly once)
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
        : "l"(p));
    return r;
}

// papr.c:103 — (I*I)+(Q*Q) with each operation rounded to float32 (the canonical x86-64 build of
// the reference emits mulss, mulss, addss; an FMA here changes the printed percentages)
// The two squares come from ONE packed instruction (Blackwell FMUL2, `mul.rn.f32x2`): each half is an IEEE
// round-to-nearest-even product with denormals kept, exactly what two FMULs give; the I/Q pair already sits in
// an aligned register pair after a 16-byte load, so nothing is moved.
__device__ __forceinline__ float power_of(float i, float q)
{
    unsigned long long p, r;
    float ii, qq;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(i), "f"(q));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(r) : "l"(p));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(ii), "=f"(qq) : "l"(r));
    return __fadd_rn(ii, qq);
}

template <int T>
__device__ __forceinline__ void track_vals(const float4 &q, float v0, float v1, float &a, float &b)
{
    if (T == TR_PEAK) { a = v0; b = v1; }
    if (T == TR_RE_POS) { a = q.x; b = q.z; }
    if (T == TR_RE_NEG) { a = -q.x; b = -q.z; }
    if (T == TR_IM_POS) { a = q.y; b = q.w; }
    if (T == TR_IM_NEG) { a = -q.y; b = -q.w; }
}

__device__ __forceinline__ double warp_sum_fixed(double x)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(FULL, x, o);
    return x; // lane 0 holds the total; same association order every run
}

// (value desc, index asc): lower index wins ties = the reference's strict compare in file order
__device__ __forceinline__ bool better(int va, u64 ia, int vb, u64 ib)
{
    return va > vb || (va == vb && ia < ib);
}

// ------------------------------------------------------------------------------------------------
// the scan kernel: pass 1 (STATS), pass 2 (HIST) or both in one sweep over the shard
// ------------------------------------------------------------------------------------------------
// dynamic shared memory of the scan kernel: u32 hist[NCELLS_MAX+2] | u32 fine_base[NCELLS_MAX+2]
// slot 0 = below the planned range, slots 1..ncells = cells, slot ncells+1 = above the range
extern __shared__ __align__(16) unsigned char scan_smem[];
#define SCAN_SMEM_BYTES (8 * (PAPR_NCELLS_MAX + 2))

template <bool STATS, bool HIST>
struct ScanState {
    // pass 1
    double dsum;
    int run_val[PAPR_NTRACK];
    unsigned run_pos[PAPR_NTRACK]; // sample offset inside this launch of the current first occurrence
    unsigned upd;
    // pass 2
    u64 *g_fine;
    unsigned smem_slot0; // shared-window byte address of slot 0 (cell c is slot c + 1)
    int sh, neg_base1, ncells; // neg_base1 = 1 - cell_base: slot = clamp((bits >> sh) + neg_base1, 0, ncells + 1)
    unsigned fmask;
};

// The CCDF pass, per sample.  No divergence up to the (warp-level infrequent) fine-table update:
// the cell index is clamped into [0, ncells+1] where slot 0 collects everything below the planned
// range (63 % of a Gaussian-like capture) and slot ncells+1 everything above it; every lane issues the
// shared-memory increment (ATOMS.POPC.INC merges lanes that hit the same word, so the crowded slot 0
// costs one update per warp) and reads the slot's fine-table base (0 = no threshold can lie here).
// SH > 0: the shift is a compile-time constant (the fused plans always use PAPR_SH_MIN), which lets shift and
// subtraction fuse into one LEA.HI; the two-sided clamp is one DPX instruction (VIMNMX.RELU).
template <bool STATS, bool HIST, int SH = 0>
__device__ __forceinline__ unsigned hist_slot(const ScanState<STATS, HIST> &st, unsigned bits)
{
    // shared-window byte address of the sample's slot
    int c = (int)(bits >> (SH > 0 ? SH : st.sh));
    // a NaN power exceeds no level (`value > level[j]` is false, papr.c:148); only the stand-alone CCDF
    // pass can meet one with levels to count against - with the statistics in the same sweep the sum is
    // NaN too and the reference prints no levels at all
    if (!STATS) c = bits > 0x7f800000u ? -1 : c; // (-1 + neg_base1 = -cell_base <= 0: slot 0 whatever the plan's base is)
    const int slot = __viaddmin_s32_relu(c, st.neg_base1, st.ncells + 1); // max(min(c + neg_base1, ncells + 1), 0)
    return st.smem_slot0 + ((unsigned)slot << 2);
}

// FB_OFF: byte distance from a cell's counter to its fine-table base (the second array of the shared window)
template <int FB_OFF = 4 * (PAPR_NCELLS_MAX + 2)>
__device__ __forceinline__ unsigned hist_bump(unsigned addr)
{
    unsigned fb;
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(fb) : "r"(addr), "n"(FB_OFF));
    return fb;
}

// one more per-value counter: fb = 1 + first fine-table counter of the sample's cell (never 0 here)
__device__ __forceinline__ void fine_bump(unsigned long long *g_fine, unsigned fb, unsigned bits, unsigned fmask)
{
    unsigned long long *p = g_fine + (fb - 1u + (bits & fmask));
    asm volatile("red.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

// the two samples of one 16-byte load
template <bool STATS, bool HIST, int FB_OFF = 4 * (PAPR_NCELLS_MAX + 2), int SH = 0>
__device__ __forceinline__ void hist_pair(ScanState<STATS, HIST> &st, float v0, float v1)
{
    const unsigned b0 = __float_as_uint(v0), b1 = __float_as_uint(v1);
    const unsigned f0 = hist_bump<FB_OFF>(hist_slot<STATS, HIST, SH>(st, b0));
    const unsigned f1 = hist_bump<FB_OFF>(hist_slot<STATS, HIST, SH>(st, b1));
    // A lane in a cell that may hold a threshold counts the sample per value.  A per-lane branch on purpose (a vote
    // would make it warp-uniform but costs more: ptxas guards every *_sync vote with a BRA.DIV), and only the lanes
    // that have such a sample enter: inside, the first update needs no predicate, and a lane with two of them (rare)
    // takes one more branch - about a third fewer instructions than two predicated updates (-g takes this path for
    // every second pair of a warp).
    if (f0 | f1) {
        const bool first = f0 != 0;
        fine_bump(st.g_fine, first ? f0 : f1, first ? b0 : b1, st.fmask);
        if (first && f1 != 0) fine_bump(st.g_fine, f1, b1, st.fmask);
    }
}

// RUNS = false: lane owns float4 u*32+lane of the batch (coalesced loads); true: float4 4*lane+u (a run
// of 8 consecutive samples per lane, papr_exact.cu)
template <int T, bool STATS, bool HIST, bool RUNS = false>
__device__ __forceinline__ void track_update(ScanState<STATS, HIST> &st, const float4 (&r)[PAPR_U],
                                             float bm, unsigned batch_off, int lane)
{
    int w = __reduce_max_sync(FULL, __float_as_int(bm));
    if (w > st.run_val[T]) { // warp-uniform and rare: locate the first sample of the batch equal to w
        unsigned pos = 0xffffffffu;
#pragma unroll
        for (int u = PAPR_U - 1; u >= 0; --u) {
            float v0 = power_of(r[u].x, r[u].y), v1 = power_of(r[u].z, r[u].w), a, b;
            track_vals<T>(r[u], v0, v1, a, b);
            unsigned p = RUNS ? 2u * (unsigned)(4 * lane + u) : 2u * (unsigned)(u * 32 + lane);
            if (__float_as_int(b) == w) pos = p + 1;
            if (__float_as_int(a) == w) pos = p;
        }
        pos = __reduce_min_sync(FULL, pos);
        st.run_val[T] = w;
        st.run_pos[T] = batch_off + pos;
        st.upd |= 1u << T;
    }
}

template <bool STATS, bool HIST>
__device__ __forceinline__ void process_batch(ScanState<STATS, HIST> &st, const float4 (&r)[PAPR_U],
                                              unsigned batch_off, int lane)
{
    float bm0 = 0.f, bm1 = 0.f, bm2 = 0.f, bm3 = 0.f, bm4 = 0.f;
#pragma unroll
    for (int u = 0; u < PAPR_U; ++u) {
        const float4 q = r[u];
        float v0 = power_of(q.x, q.y);
        float v1 = power_of(q.z, q.w);
        if (STATS) {
            st.dsum += (double)v0; // papr.c:104 (order differs from the file order)
            st.dsum += (double)v1;
            // fmaxf drops NaN operands, like the reference's comparisons (papr.c:105-126)
            bm0 = fmaxf(bm0, fmaxf(v0, v1));
            bm1 = fmaxf(bm1, fmaxf(q.x, q.z));
            bm2 = fmaxf(bm2, fmaxf(-q.x, -q.z));
            bm3 = fmaxf(bm3, fmaxf(q.y, q.w));
            bm4 = fmaxf(bm4, fmaxf(-q.y, -q.w));
        }
        if (HIST) hist_pair(st, v0, v1);
    }
    if (STATS) {
        // warp-uniform and rare after the first few batches: does any lane beat a running maximum?
        const bool cand = __float_as_int(bm0) > st.run_val[TR_PEAK] || __float_as_int(bm1) > st.run_val[TR_RE_POS] ||
                          __float_as_int(bm2) > st.run_val[TR_RE_NEG] || __float_as_int(bm3) > st.run_val[TR_IM_POS] ||
                          __float_as_int(bm4) > st.run_val[TR_IM_NEG];
        if (__any_sync(FULL, cand)) {
            track_update<TR_PEAK>(st, r, bm0, batch_off, lane);
            track_update<TR_RE_POS>(st, r, bm1, batch_off, lane);
            track_update<TR_RE_NEG>(st, r, bm2, batch_off, lane);
            track_update<TR_IM_POS>(st, r, bm3, batch_off, lane);
            track_update<TR_IM_NEG>(st, r, bm4, batch_off, lane);
        }
    }
}

