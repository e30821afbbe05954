// papr_tr.cu — tone-reservation PAPR reduction of OFDM symbols.
//
// The reference configures GNU Radio's `dtv.dvbt2_paprtr_cc(..., vclip = 3.3, iterations = 3, fftsize)` in its
// DVB-T2 transmit chain (drmpeg/dtv-utils dvbt2-blade.py:52-54,129) and budgets the reserved cells in
// dvbt2rate.c:1217; the block itself lives in gr-dtv, outside the reference tree.  What it implements is the
// iterative gradient algorithm of EN 302 755 clause 9.6.2.1, restated here for the time-domain symbols:
//     x      the symbol after the IFFT (N samples), corrected in place
//     p      the reference kernel: IFFT of "1 on every reserved tone", scaled so that p[0] = 1
//     r_k    what the reserved tone k carries so far (starts at 0; |r_k| must stay <= Amax)
//   per iteration: y = max |x_n| at n = m (first maximum); stop if y <= Vclip (+ 4 ppm: float32); u = x_m / y;
//     v_k = u e^{-j 2 pi k m / N};  alpha_k = sqrt(Amax^2 - Im(r_k conj v_k)^2) + Re(r_k conj v_k);
//     alpha = min(y - Vclip, min_k alpha_k); stop if alpha <= 0;
//     x_n -= alpha u p[(n - m) mod N];  r_k -= alpha v_k.
// One CTA per symbol, the symbol stays in L2 (256 KB at 32K) between the three iterations: each iteration is an
// arg-max sweep and an update sweep over the symbol - HBM/L2-bound elementwise work, no tensor cores.
// The evaluation of what it buys is the CCDF engine itself (papr_analyze_device before / after).
#include <cuda_runtime.h>
#include <math.h>

#include "papr_device.cuh"

#define TR_T 1024

__global__ void __launch_bounds__(TR_T) papr_tr_kernel(float2 *x, int nsym, int N, const float2 *p, const int *tone, int ntones,
                                                      float vclip, int iterations, float amax, float2 *r_out, int *iters_out)
{
    __shared__ float s_v[TR_T / 32];
    __shared__ int s_i[TR_T / 32];
    __shared__ float s_a[TR_T / 32];
    __shared__ float s_y, s_alpha;
    __shared__ int s_m;
    __shared__ float2 s_u;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int sym = blockIdx.x; sym < nsym; sym += gridDim.x) {
        float2 *xs = x + (size_t)sym * N;
        float2 *rs = r_out + (size_t)sym * ntones;
        for (int k = t; k < ntones; k += TR_T) rs[k] = make_float2(0.f, 0.f);
        int it = 0;
        for (; it < iterations; ++it) {
            // ---- y = max |x_n|, m = its first index (squared magnitudes; three separately rounded operations)
            float best = -1.f;
            int bi = 0x7fffffff;
            for (int n = t; n < N; n += TR_T) {
                const float2 q = xs[n];
                const float v = __fadd_rn(__fmul_rn(q.x, q.x), __fmul_rn(q.y, q.y));
                if (v > best) { best = v; bi = n; } // ascending n per thread: the first maximum is kept
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_v[warp] = best; s_i[warp] = bi; }
            __syncthreads();
            if (t == 0) {
                float b = s_v[0];
                int i = s_i[0];
                for (int w = 1; w < TR_T / 32; ++w)
                    if (s_v[w] > b || (s_v[w] == b && s_i[w] < i)) { b = s_v[w]; i = s_i[w]; }
                const float2 q = xs[i];
                const float y = sqrtf(b);
                s_y = y; s_m = i;
                s_u = y > 0.f ? make_float2(q.x / y, q.y / y) : make_float2(0.f, 0.f);
            }
            __syncthreads();
            const float y = s_y;
            const int m = s_m;
            const float2 u = s_u;
            // (uniform) a peak already brought down to Vclip sits there to within float32 rounding: not a new peak
            if (!(y > vclip * (1.f + 4e-6f))) break;
            // ---- the largest step every reserved tone still allows
            float amin = y - vclip;
            for (int k = t; k < ntones; k += TR_T) {
                float sn, cs;
                sincospif(-2.f * (float)(((long long)tone[k] * m) % N) / (float)N, &sn, &cs);
                const float2 v = make_float2(u.x * cs - u.y * sn, u.x * sn + u.y * cs);
                const float2 r = rs[k];
                const float re = r.x * v.x + r.y * v.y, im = r.y * v.x - r.x * v.y; // r conj(v)
                const float a = sqrtf(fmaxf(amax * amax - im * im, 0.f)) + re;
                amin = fminf(amin, a);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amin = fminf(amin, __shfl_down_sync(0xffffffffu, amin, o));
            if (lane == 0) s_a[warp] = amin;
            __syncthreads();
            if (t == 0) {
                float a = s_a[0];
                for (int w = 1; w < TR_T / 32; ++w) a = fminf(a, s_a[w]);
                s_alpha = a;
            }
            __syncthreads();
            const float alpha = s_alpha;
            if (!(alpha > 0.f)) break;
            // ---- x_n -= alpha u p[(n - m) mod N];  r_k -= alpha v_k
            const float2 au = make_float2(alpha * u.x, alpha * u.y);
            for (int n = t; n < N; n += TR_T) {
                int j = n - m;
                if (j < 0) j += N;
                const float2 pk = p[j];
                float2 q = xs[n];
                q.x -= au.x * pk.x - au.y * pk.y;
                q.y -= au.x * pk.y + au.y * pk.x;
                xs[n] = q;
            }
            for (int k = t; k < ntones; k += TR_T) {
                float sn, cs;
                sincospif(-2.f * (float)(((long long)tone[k] * m) % N) / (float)N, &sn, &cs);
                float2 r = rs[k];
                r.x -= au.x * cs - au.y * sn;
                r.y -= au.x * sn + au.y * cs;
                rs[k] = r;
            }
            __syncthreads();
        }
        if (t == 0 && iters_out) iters_out[sym] = it;
        __syncthreads();
    }
}

void papr_launch_tr(float *x, int nsym, int N, const float *p, const int *tone, int ntones, float vclip, int iterations,
                    float amax, float *r_out, int *iters_out, int grid, cudaStream_t s)
{
    papr_tr_kernel<<<grid, TR_T, 0, s>>>(reinterpret_cast<float2 *>(x), nsym, N, reinterpret_cast<const float2 *>(p), tone,
                                         ntones, vclip, iterations, amax, reinterpret_cast<float2 *>(r_out), iters_out);
}
