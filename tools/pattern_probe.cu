// tools/pattern_probe.cu — which global->register/shared access patterns give every LANE a run of
// CONSECUTIVE samples (needed to fold the sequential-sum emulation into the scan) and still read HBM
// at the ceiling of the part?  Developer probe, not part of the product.
//   cyc   : batch-cyclic work split of the shipped scan kernel (warp gw takes batches gw, gw+GW, ...)
//   tiled : every warp streams its own 32 KiB tile, a CTA owns a contiguous 1 MiB super-tile
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

struct f8 { float v[8]; };

__device__ __forceinline__ float4 ld128(const void *p, bool alloc)
{
    float4 r;
    if (alloc)
        asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ f8 ld256(const void *p)
{
    f8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}

// batch index sequence of one warp
template <bool TILED>
struct Walk {
    size_t b, end, nb, GW;
    size_t tile, ntiles, cta_stride;
    int warp;
    static constexpr int TB = 16; // batches per tile (2 KiB batches -> 32 KiB tiles)
    __device__ Walk(size_t nbatch, int bpw /* batch bytes / 2048 */)
    {
        nb = nbatch;
        warp = threadIdx.x >> 5;
        if (!TILED) {
            GW = (size_t)gridDim.x * (blockDim.x >> 5);
            b = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
            end = nb;
        } else {
            const int tb = TB / bpw;
            ntiles = nb / tb;
            tile = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
            cta_stride = (size_t)gridDim.x * (blockDim.x >> 5);
            b = tile * tb;
            end = tile < ntiles ? b + tb : b;
            GW = 1;
        }
    }
    __device__ bool next(int bpw)
    {
        b += GW;
        if (b < end) return true;
        if (!TILED) return false;
        const int tb = TB / bpw;
        tile += cta_stride;
        if (tile >= ntiles) return false;
        b = tile * tb;
        end = b + tb;
        return true;
    }
};

// MODE 0: coalesced 4 x LDG.128 (shipped)                      lane run = 2 samples
// MODE 1: 4 x LDG.128 at lane stride 64 B, L1 no_allocate       lane run = 8 samples
// MODE 2: same, L1 allocating                                   lane run = 8 samples
// MODE 3: 2 x LDG.256 at lane stride 64 B                       lane run = 8 samples
// MODE 4: 4 x LDG.256 at lane stride 128 B (4 KiB batches)      lane run = 16 samples
// MODE 5: 2 x LDG.256 coalesced (lane stride 32 B), 2 KiB batch lane run = 4 samples (x2)
template <int MODE, bool TILED>
__global__ void __launch_bounds__(1024) read_direct(const char *p, size_t nbytes, float *out)
{
    const int lane = threadIdx.x & 31;
    constexpr int BB = MODE == 4 ? 4096 : 2048;
    Walk<TILED> w(nbytes / BB, BB / 2048);
    float acc = 0.f;
    if (w.b >= w.end) return;
    do {
        const char *q = p + w.b * BB;
        if (MODE == 0) {
            float4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = ld128(q + 16 * (32 * u + lane), false);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += r[u].x + r[u].y + r[u].z + r[u].w;
        } else if (MODE == 1 || MODE == 2) {
            float4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = ld128(q + 64 * lane + 16 * u, MODE == 2);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc += r[u].x + r[u].y + r[u].z + r[u].w;
        } else if (MODE == 3) {
            f8 r[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) r[u] = ld256(q + 64 * lane + 32 * u);
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += r[u].v[k];
        } else if (MODE == 4) {
            f8 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = ld256(q + 128 * lane + 32 * u);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += r[u].v[k];
        } else {
            f8 r[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) r[u] = ld256(q + 32 * (32 * u + lane));
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc += r[u].v[k];
        }
    } while (w.next(BB / 2048));
    if (acc == 123.456f) out[0] = acc;
}

// cp.async (LDGSTS): coalesced 16-byte global reads written straight into a swizzled shared layout,
// then every lane reads its own 64-byte run (4 x LDS.128, conflict-free)
template <bool TILED>
__global__ void __launch_bounds__(1024) read_ldgsts(const char *p, size_t nbytes, float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *my = smem + warp * 2048;
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(my);
    Walk<TILED> w(nbytes / 2048, 1);
    float acc = 0.f;
    if (w.b >= w.end) return;
    // chunk index c = 16-byte piece of the batch; physical = 8a + (b ^ (a & 7)), c = 8a + b
    unsigned st_off[4], ld_off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int c = 32 * u + lane, a = c >> 3, b = c & 7;
        st_off[u] = 16u * (8 * a + (b ^ (a & 7)));
        c = 4 * lane + u; a = c >> 3; b = c & 7;
        ld_off[u] = 16u * (8 * a + (b ^ (a & 7)));
    }
    do {
        const char *q = p + w.b * 2048;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm0 + st_off[u]), "l"(q + 16 * (32 * u + lane)) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 r = *reinterpret_cast<const float4 *>(my + ld_off[u]);
            acc += r.x + r.y + r.z + r.w;
        }
        __syncwarp();
    } while (w.next(1));
    if (acc == 123.456f) out[0] = acc;
}

// TMA 2-D tile [ROWS x 128 B] with the 128-byte swizzle, one single-stage buffer per warp; every lane
// then reads its own run (ROWS=16: half a row = 64 B; ROWS=32: a whole row = 128 B) with LDS.128
template <int ROWS, bool TILED>
__global__ void __launch_bounds__(1024) read_tma(const __grid_constant__ CUtensorMap tmap, size_t nbytes, float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int BB = ROWS * 128;
    unsigned char *my = smem + warp * BB;
    const unsigned sm0 = (unsigned)__cvta_generic_to_shared(my);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bars[warp]);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    Walk<TILED> w(nbytes / BB, BB / 2048);
    float acc = 0.f;
    if (w.b >= w.end) return;
    auto issue = [&](size_t b) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BB) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(sm0), "l"(&tmap), "r"(0), "r"((int)(b * ROWS)), "r"(bar) : "memory");
    };
    unsigned off[8];
    constexpr int NL = ROWS == 16 ? 4 : 8;
#pragma unroll
    for (int j = 0; j < NL; ++j) {
        const int row = ROWS == 16 ? lane >> 1 : lane;
        const int c = ROWS == 16 ? 4 * (lane & 1) + j : j;
        off[j] = row * 128 + ((c ^ (row & 7)) << 4);
    }
    if (lane == 0) issue(w.b);
    unsigned phase = 0;
    for (;;) {
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        phase ^= 1;
        float4 r[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) r[j] = *reinterpret_cast<const float4 *>(my + off[j]);
        __syncwarp();
        const bool more = w.next(BB / 2048);
        if (more && lane == 0) issue(w.b);
#pragma unroll
        for (int j = 0; j < NL; ++j) acc += r[j].x + r[j].y + r[j].z + r[j].w;
        if (!more) break;
    }
    if (acc == 123.456f) out[0] = acc;
}

template <typename F> float timeit(F f)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9, ms;
    for (int i = 0; i < 6; ++i) { cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1); if (i >= 2 && ms < best) best = ms; }
    return best;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const size_t bytes = 16ull << 30;
    char *d; float *o; CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&o, 4)); CK(cudaMemset(d, 1, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char *n, float ms) { printf("%-46s %.3f ms  %.0f GB/s\n", n, ms, bytes / ms / 1e6); fflush(stdout); };
#define DIRECT(M, T, name) rep(name, timeit([&] { read_direct<M, T><<<sms, 1024>>>(d, bytes, o); }))
    DIRECT(0, false, "cyc   4xLDG.128 coalesced (shipped)");
    DIRECT(0, true,  "tiled 4xLDG.128 coalesced");
    DIRECT(1, false, "cyc   4xLDG.128 lane-stride 64B noalloc");
    DIRECT(1, true,  "tiled 4xLDG.128 lane-stride 64B noalloc");
    DIRECT(2, false, "cyc   4xLDG.128 lane-stride 64B L1 alloc");
    DIRECT(2, true,  "tiled 4xLDG.128 lane-stride 64B L1 alloc");
    DIRECT(3, false, "cyc   2xLDG.256 lane-stride 64B");
    DIRECT(3, true,  "tiled 2xLDG.256 lane-stride 64B");
    DIRECT(4, false, "cyc   4xLDG.256 lane-stride 128B (4K batch)");
    DIRECT(4, true,  "tiled 4xLDG.256 lane-stride 128B (4K batch)");
    DIRECT(5, false, "cyc   2xLDG.256 coalesced");
    DIRECT(5, true,  "tiled 2xLDG.256 coalesced");
    {
        auto k0 = read_ldgsts<false>; auto k1 = read_ldgsts<true>;
        CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        rep("cyc   4xLDGSTS swizzled + 4xLDS.128 runs", timeit([&] { k0<<<sms, 1024, 65536>>>(d, bytes, o); }));
        rep("tiled 4xLDGSTS swizzled + 4xLDS.128 runs", timeit([&] { k1<<<sms, 1024, 65536>>>(d, bytes, o); }));
    }
    {
        EncodeFn enc = nullptr;
        cudaDriverEntryPointQueryResult qr;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr));
        if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
        for (int rows = 16; rows <= 32; rows += 16) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {32, bytes / 128};
            cuuint64_t strides[1] = {128};
            cuuint32_t box[2] = {32, (cuuint32_t)rows};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            const int sm = 32 * rows * 128;
            if (rows == 16) {
                auto k0 = read_tma<16, false>; auto k1 = read_tma<16, true>;
                CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                rep("cyc   TMA 2D 16x128B swz128 + 4xLDS.128 runs", timeit([&] { k0<<<sms, 1024, sm>>>(tm, bytes, o); }));
                rep("tiled TMA 2D 16x128B swz128 + 4xLDS.128 runs", timeit([&] { k1<<<sms, 1024, sm>>>(tm, bytes, o); }));
            } else {
                auto k0 = read_tma<32, false>; auto k1 = read_tma<32, true>;
                CK(cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                rep("cyc   TMA 2D 32x128B swz128 + 8xLDS.128 runs", timeit([&] { k0<<<sms, 1024, sm>>>(tm, bytes, o); }));
                rep("tiled TMA 2D 32x128B swz128 + 8xLDS.128 runs", timeit([&] { k1<<<sms, 1024, sm>>>(tm, bytes, o); }));
            }
        }
    }
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}
