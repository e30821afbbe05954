/*
 * oracle/papr_oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference tool's algorithm (drmpeg/dtv-utils papr.c:32-196) used only
 * as the parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 * Nothing under dtv-utils_b200/ may include, link or call this.
 *
 * Parity status: PINNED.  The restatement is checked byte-for-byte on stdout against the
 * unmodified reference binary (oracle/_ref/papr, built by oracle/Makefile from
 * /root/reference/papr.c) on every fixture in tests/golden/ (see tests/test_oracle.py and
 * tests/golden/make_golden.py).  The reference ships no tests or golden vectors of its own
 * (SURVEY.md §4.1), so outputs of the reference itself are the pin.
 */
#ifndef PAPR_ORACLE_H
#define PAPR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAPR_ORACLE_CHUNK 16384 /* floats per fread, papr.c:30 */

/* pass-1 state, papr.c:36-49 */
typedef struct {
    int64_t n;            /* offset            papr.c:37,127 */
    double  sum;          /* sum               papr.c:39,104 (sequential) */
    float   peak;         /* peak              papr.c:40,105-108 */
    int64_t peak_idx;     /* peak_offset */
    float   re_pos, im_pos, re_neg, im_neg;                 /* papr.c:42-45 */
    int64_t re_pos_idx, im_pos_idx, re_neg_idx, im_neg_idx; /* papr.c:46-49 */
} papr_oracle_stats;

void papr_oracle_stats_init(papr_oracle_stats *st);

/* papr.c:100-129 over `nsamples` complete I/Q pairs, continuing a running state (so a caller can
 * feed a file piecewise, or append the odd-tail sample). */
void papr_oracle_pass1(const float *iq, int64_t nsamples, papr_oracle_stats *st);

/* papr.c:131,134,136-141 (graph==0) / 164-173 (graph!=0).  Returns L = number of levels
 * (0 when papr is NaN or negative enough that the reference's loops do not run), writes
 * avg (= sum/offset), papr, and level[0..min(L,cap)).  */
int papr_oracle_levels(const papr_oracle_stats *st, int graph, double *avg, float *papr,
                       float *level, int cap);

/* papr.c:143-153 / 175-185: level_count[j] += #{samples with power > level[j]}.
 * Histogram + suffix sum when level[] is non-decreasing (proved equivalent);
 * the literal O(N*L) loop otherwise. */
void papr_oracle_pass2(const float *iq, int64_t nsamples, const float *level, int L,
                       int64_t *level_count);

/* papr.c:132-135,154-161 / 186-190: the exact stdout text.  Returns bytes written (excluding
 * the terminating NUL) or -1 if cap is too small. */
long papr_oracle_format(const papr_oracle_stats *st, int graph, double avg, float papr,
                        const int64_t *level_count, int L, char *out, size_t cap);

/* Whole tool on a byte buffer holding the file image (emulates the 64 KiB fread chunking,
 * including the stale-Q behaviour on files with an odd number of floats, papr.c:100-103). */
long papr_oracle_run_buffer(const void *file_image, size_t file_bytes, int graph, char *out,
                            size_t cap);
long papr_oracle_run_file(const char *path, int graph, char *out, size_t cap);

/* integer-only counter-based synthetic I/Q generator (C twin of the CUDA
 * generator in the product library).  Writes 2*nsamples floats for samples
 * [first, first+nsamples). */
void papr_oracle_siggen(float *iq, uint64_t first, uint64_t nsamples, uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
