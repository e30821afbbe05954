// tools/scan_tune.cu — developer harness: times the scan kernel variants selected with -D macros
// (PAPR_THREADS, PAPR_U, PAPR_CTAS_PER_SM) on a device-generated capture.  Not part of the product.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../dtv-utils_b200/csrc/papr_kernels.cu"
extern "C" void papr_host_build_tables(int graph, int n, double *pow10, double *ratio_min);

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

int main(int argc, char **argv)
{
    int log2n = argc > 1 ? atoi(argv[1]) : 31;
    u64 n = 1ull << log2n;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int grid = sms * PAPR_CTAS_PER_SM, nwarps = grid;
    float *iq; CK(cudaMalloc(&iq, n * 8));
    papr_launch_siggen(iq, 0, n, 1, sms * 16, 0);
    PaprCtaPartial *wp; CK(cudaMalloc(&wp, sizeof(PaprCtaPartial) * nwarps));
    u64 *hist, *fine, *over; unsigned *fb; PaprPlan *plan; double *prewp, *pre4, *tab;
    CK(cudaMalloc(&hist, 8 * PAPR_NCELLS_MAX)); CK(cudaMalloc(&over, 8)); CK(cudaMalloc(&fine, 64 << 20));
    CK(cudaMalloc(&fb, 4 * PAPR_NCELLS_MAX)); CK(cudaMalloc(&plan, sizeof(PaprPlan)));
    CK(cudaMalloc(&prewp, 8 * 3 * nwarps)); CK(cudaMalloc(&pre4, 32)); CK(cudaMalloc(&tab, 8 * 4 * PAPR_MAX_LEVELS));
    std::vector<double> t(4 * PAPR_MAX_LEVELS, INFINITY);
    papr_host_build_tables(0, 256, &t[0], &t[2 * PAPR_MAX_LEVELS]);
    papr_host_build_tables(1, PAPR_MAX_LEVELS, &t[PAPR_MAX_LEVELS], &t[3 * PAPR_MAX_LEVELS]);
    CK(cudaMemcpy(tab, t.data(), 8 * t.size(), cudaMemcpyHostToDevice));
    papr_scan_configure();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("threads=%d U=%d ctas/sm=%d grid=%d n=2^%d\n", PAPR_THREADS, PAPR_U, PAPR_CTAS_PER_SM, grid, log2n);
    for (int graph = 0; graph < 2; ++graph) {
        PaprTables tb; tb.pow10 = tab + (graph ? PAPR_MAX_LEVELS : 0); tb.ratio_min = tab + 2 * PAPR_MAX_LEVELS + (graph ? PAPR_MAX_LEVELS : 0);
        tb.nlevels_max = graph ? PAPR_MAX_LEVELS : 256;
        int stride = graph ? 32 : 128;
        papr_launch_presample(iq, n, stride, grid, prewp, 0);
        papr_launch_plan_pred(nullptr, prewp, grid, tb, 5.0f, 1.0f, 2048, plan, fb, 0);
        PaprPlan hp; CK(cudaMemcpy(&hp, plan, sizeof(hp), cudaMemcpyDeviceToHost));
        printf(" graph=%d plan: sh=%d base=%d ncells=%d n_amb=%d cov=%d w=%g\n", graph, hp.sh, hp.cell_base, hp.ncells, hp.n_amb, hp.levels_covered, hp.window);
        PaprScanArgs a; a.iq = iq; a.nsamples = n; a.first_index = 0; a.wp = wp; a.plan = plan; a.fine_base = fb; a.g_hist = hist; a.g_fine = fine; a.g_over = over;
        for (int variant = 0; variant < 3; ++variant) {
            bool st = variant != 1, hi = variant != 0;
            float best = 1e9, ms;
            for (int it = 0; it < 6; ++it) {
                cudaMemsetAsync(wp, 0, sizeof(PaprCtaPartial) * nwarps, 0);
                cudaMemsetAsync(hist, 0, 8 * PAPR_NCELLS_MAX, 0);
                papr_launch_zero_fine(plan, fine, sms * 4, 0);
                cudaEventRecord(e0, 0);
                papr_launch_scan(st, hi, grid, a, 0);
                cudaEventRecord(e1, 0);
                CK(cudaEventSynchronize(e1));
                cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 2 && ms < best) best = ms;
            }
            printf("  scan<stats=%d,hist=%d>  %.3f ms  %.0f GB/s\n", st, hi, best, 8.0 * n / best / 1e6);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
