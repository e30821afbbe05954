// papr_xchg.cuh — the in-kernel exchange between the ranks of one box: publish into every peer's window over
// NVLink (cudaIpc-mapped), release-store a flag, acquire-poll the own window.  Shared by papr_kernels.cu (plan /
// statistics / counts exchanges) and papr_exact.cu (the chain of the sequential sum).  PaprXchg: papr_device.cuh.
#pragma once
#include "papr_device.cuh"

typedef unsigned long long u64;

__device__ __forceinline__ void st_release_sys(u64 *p, u64 v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u64 ld_acquire_sys(const u64 *p)
{
    u64 v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ u64 ld_volatile(const u64 *p)
{
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ u64 global_timer_ns()
{
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All threads of ONE CTA: copy `words` u64 from src into the field at `field_off` of slot[rank] in
// every rank's window (NVLink stores for the peers, a local store for the own window), then raise
// flag[kind][rank] = seq everywhere with a system-scope release.
static __device__ void xchg_publish(const PaprPeers &pp, int kind, size_t field_off, const u64 *src, int words, u64 seq)
{
    for (int r = 0; r < pp.world; ++r) {
        u64 *dst = reinterpret_cast<u64 *>(reinterpret_cast<char *>(&pp.win[r]->slot[pp.rank]) + field_off);
        for (int i = threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pp.world) st_release_sys(&pp.win[threadIdx.x]->flag[kind][pp.rank], seq);
}

// All threads of ONE CTA: wait until every rank's publication `seq` of `kind` has landed in the own
// window.  Returns false if a peer did not show up within pp.timeout_ns (tunable "xchg_timeout_s") or
// some rank already gave up: the rank that times out raises its abort word in EVERY window before it
// returns, so a peer that arrives late finds it and fails as well - all ranks agree on the outcome.
static __device__ bool xchg_wait(const PaprPeers &pp, int kind, u64 seq)
{
    __shared__ int s_late;
    if (threadIdx.x == 0) s_late = 0;
    __syncthreads();
    if ((int)threadIdx.x < pp.world) {
        const u64 *f = &pp.win[pp.rank]->flag[kind][threadIdx.x];
        const u64 *ab = &pp.win[pp.rank]->abort[threadIdx.x];
        const u64 t0 = global_timer_ns();
        while (ld_acquire_sys(f) < seq) {
            if (ld_volatile(ab) != 0 || global_timer_ns() - t0 > pp.timeout_ns) { atomicExch(&s_late, 1); break; }
            __nanosleep(200);
        }
        if (ld_volatile(ab) != 0) atomicExch(&s_late, 1);
    }
    __syncthreads();
    const bool late = s_late != 0;
    if (late && (int)threadIdx.x < pp.world) st_release_sys(&pp.win[threadIdx.x]->abort[pp.rank], seq);
    return !late;
}


// All threads of ONE CTA: publish this shard's level counts (+ status word), collect every rank's, add them up in
// place.  Only the L levels in use (and the status word) travel; buffer seq & 1 of the slot is written, so a second
// publication within one analysis (recount / exact redo) cannot race with a peer that is still summing the first.
// `counts` may have been written by other CTAs of the calling kernel (after a __threadfence + ticket): read past L1.
static __device__ void xchg_counts(u64 *counts, int L, PaprPlan *plan, const PaprPeers &pp, u64 seq)
{
    const int buf = (int)(seq & 1);
    const size_t off = offsetof(PaprXchgSlot, counts) + (size_t)buf * sizeof(u64) * (PAPR_MAX_LEVELS + 1);
    for (int r = 0; r < pp.world; ++r) {
        u64 *dst = reinterpret_cast<u64 *>(reinterpret_cast<char *>(&pp.win[r]->slot[pp.rank]) + off);
        for (int i = threadIdx.x; i < L; i += blockDim.x) dst[i] = ld_volatile(&counts[i]);
        if (threadIdx.x == 0) dst[PAPR_MAX_LEVELS] = ld_volatile(&counts[PAPR_MAX_LEVELS]);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pp.world) st_release_sys(&pp.win[threadIdx.x]->flag[XK_COUNTS][pp.rank], seq);
    const bool ok = xchg_wait(pp, XK_COUNTS, seq);
    if (!ok && threadIdx.x == 0) plan->pad = 1;
    for (int i = threadIdx.x; i <= PAPR_MAX_LEVELS; i += blockDim.x) {
        if (i >= L && i != PAPR_MAX_LEVELS) continue;
        u64 acc = 0;
        for (int q = 0; q < pp.world; ++q) acc += ld_volatile(&pp.win[pp.rank]->slot[q].counts[buf][i]);
        counts[i] = acc;
    }
}
