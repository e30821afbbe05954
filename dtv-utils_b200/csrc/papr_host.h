/* papr_host.h — internal host-side helpers shared by papr_host.c, papr_engine.cu and papr_main.c */
#ifndef PAPR_HOST_H
#define PAPR_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* pow10[j] = pow(10, x_j) and ratio_min[j] = least peak/avg for which the reference reaches level j,
 * both with the host libm, for j < n (graph selects papr.c:139 vs papr.c:168-172 exponents) */
void papr_host_build_tables(int graph, int n, double *pow10, double *ratio_min);
/* Q that the reference pairs with a lone trailing I (papr.c:35,101-103) */
float papr_host_stale_q(const unsigned char *file_image, uint64_t file_bytes);
/* sum += (double)(I*I+Q*Q) sample by sample, exactly as papr.c:103-104 */
double papr_host_seq_add(double sum, const float *iq, uint64_t nsamples);
#ifdef __cplusplus
}
#endif
#endif
