/*
 * papr_host.c — the scalar epilogue and the text output of the reference tool, restated for the
 * product library (host side, plain C, same glibc libm/printf as the reference binary uses).
 *
 * Replaces drmpeg/dtv-utils papr.c:131-141 (1 dB levels), :164-173 (-g levels), :132-135,154-161 and
 * :186-190 (printf).  These few hundred scalar operations decide every printed digit, so they
 * follow the reference's C conversion sequence exactly and must be compiled
 * without -ffast-math / -march=native and with -ffp-contract=off.
 *
 * Also builds, once per engine, the two tables that let the GPU evaluate the same epilogue with
 * only IEEE multiply/divide/compare (papr_kernels.cu: papr_levels_kernel).
 */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/papr_b200.h"
#include "papr_host.h"

/* (int)x as the reference's x86-64 build evaluates it: cvttss2si yields INT_MIN for NaN and for
 * values outside int range, which makes the reference's `i <= (int)papr` loops not run. */
static int int_of_float(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (-2147483647 - 1);
    return (int)x;
}

/* highest level index the reference's loops reach: (int)papr (papr.c:136-138) or (int)(papr*10)
 * (papr.c:166-169, a float multiply) */
static int top_level(float papr, int graph)
{
    if (!graph) return int_of_float(papr);
    volatile float p10 = papr * 10.0f;
    return int_of_float(p10);
}

/* papr.c:134 / 165: papr = 10 * log10(peak / sum), evaluated in double, stored to float */
static float papr_of_ratio(double ratio)
{
    return (float)(10.0 * log10(ratio));
}

/* exponent argument of level j: (float)j / 10 (papr.c:139) or the float-accumulated index / 10
 * with index += 0.1 evaluated in double and rounded back to float each step (papr.c:168-172) */
static void level_exponents(int graph, int n, float *x)
{
    if (!graph) {
        for (int j = 0; j < n; j++) {
            volatile float v = (float)j / 10.0f;
            x[j] = v;
        }
    } else {
        float index = 0.0f;
        for (int j = 0; j < n; j++) {
            volatile float v = index / 10.0f;
            x[j] = v;
            index = (float)((double)index + 0.1);
        }
    }
}

void papr_host_build_tables(int graph, int n, double *pow10, double *ratio_min)
{
    static float x[PAPR_MAX_LEVELS];
    if (n > PAPR_MAX_LEVELS) n = PAPR_MAX_LEVELS;
    level_exponents(graph, n, x);
    for (int j = 0; j < n; j++) pow10[j] = pow(10.0, (double)x[j]);
    /* ratio_min[j] = least positive double r with top_level(papr_of_ratio(r)) >= j, found by
     * bisection on the bit pattern (positive doubles order like their bits).  Relies on log10 being
     * monotone; the engine re-derives L on the host from the final ratio and compares. */
    for (int j = 0; j < n; j++) {
        unsigned long long lo = 1ull, hi = 0x7ff0000000000000ull; /* (min denormal, +inf] */
        while (lo < hi) {
            unsigned long long mid = lo + ((hi - lo) >> 1);
            double r;
            memcpy(&r, &mid, 8);
            if (top_level(papr_of_ratio(r), graph) >= j) hi = mid; else lo = mid + 1;
        }
        memcpy(&ratio_min[j], &lo, 8);
    }
}

int papr_levels(const papr_stats *st, int graph, double *avg_out, float *papr_out, float *level, int cap)
{
    volatile double avg = st->sum / (double)(long long)st->n; /* papr.c:131 / 164 */
    volatile double ratio = (double)st->peak / avg;
    float papr = papr_of_ratio(ratio);                        /* papr.c:134 / 165 */
    if (avg_out) *avg_out = avg;
    if (papr_out) *papr_out = papr;
    int top = top_level(papr, graph);
    if (top < 0) return 0;
    int L = top + 1;
    if (L > PAPR_MAX_LEVELS) L = PAPR_MAX_LEVELS; /* unreachable: PAPR < 192.7 dB, see papr_b200.h */
    if (level && cap > 0) {
        static __thread float x[PAPR_MAX_LEVELS];
        int n = L < cap ? L : cap;
        level_exponents(graph, n, x);
        for (int j = 0; j < n; j++) level[j] = (float)(pow(10.0, (double)x[j]) * avg); /* :139 / :170 */
    }
    return L;
}

void papr_stats_merge(papr_stats *a, const papr_stats *b)
{
    /* `b` covers the samples right after `a`: strict compares keep the earlier first occurrence */
    a->sum += b->sum;
    a->n += b->n;
    if (b->peak > a->peak) { a->peak = b->peak; a->peak_idx = b->peak_idx; }
    if (b->re_pos > a->re_pos) { a->re_pos = b->re_pos; a->re_pos_idx = b->re_pos_idx; }
    if (b->re_neg < a->re_neg) { a->re_neg = b->re_neg; a->re_neg_idx = b->re_neg_idx; }
    if (b->im_pos > a->im_pos) { a->im_pos = b->im_pos; a->im_pos_idx = b->im_pos_idx; }
    if (b->im_neg < a->im_neg) { a->im_neg = b->im_neg; a->im_neg_idx = b->im_neg_idx; }
    a->flags |= b->flags;
}

int papr_result_finish(papr_result *r, int graph)
{
    r->graph = graph;
    r->nlevels = papr_levels(&r->stats, graph, &r->avg, &r->papr, r->level, PAPR_MAX_LEVELS);
    return r->nlevels;
}

#define EMIT(...)                                                     \
    do {                                                              \
        int w_ = snprintf(out + pos, cap - (size_t)pos, __VA_ARGS__); \
        if (w_ < 0 || (size_t)w_ >= cap - (size_t)pos) return -1;     \
        pos += w_;                                                    \
    } while (0)

long papr_format(const papr_result *r, char *out, size_t cap)
{
    const papr_stats *st = &r->stats;
    long pos = 0;
    if (!out || cap == 0) return -1;
    out[0] = 0;
    if (!r->graph) {
        EMIT("Peak magnitude = %f\n", sqrt((double)st->peak));                                   /* :132 */
        EMIT("average power = %lf, peak power = %f @ %lld\n\n", r->avg, (double)st->peak,
             (long long)st->peak_idx * 8);                                                        /* :133 */
        EMIT("Maximum PAPR = %f\n", (double)r->papr);                                             /* :135 */
        for (int i = 0; i < r->nlevels; i++) {                                                    /* :154-156 */
            volatile float frac = (float)(long long)r->level_count[i] / (float)(long long)st->n;
            EMIT("percentage above %d dB = %0.8f\n", i, (double)frac * 100.0);
        }
        EMIT("\n");                                                                               /* :157 */
        EMIT("peak real positive = %f, peak imaginary positive = %f\n", (double)st->re_pos,
             (double)st->im_pos);                                                                 /* :158 */
        EMIT("peak real negative = %f, peak imaginary negative = %f\n\n", (double)st->re_neg,
             (double)st->im_neg);                                                                 /* :159 */
        EMIT("peak real positive @ %lld, peak imaginary positive @ %lld\n",
             (long long)st->re_pos_idx * 8, (long long)st->im_pos_idx * 8 + 1);                   /* :160 */
        EMIT("peak real negative @ %lld, peak imaginary negative @ %lld\n",
             (long long)st->re_neg_idx * 8, (long long)st->im_neg_idx * 8 + 1);                   /* :161 */
    } else {
        for (int i = 0; i < r->nlevels; i++) {                                                    /* :187-190 */
            volatile float frac = (float)(long long)r->level_count[i] / (float)(long long)st->n;
            EMIT("%0.8f\n", (double)frac * 100.0);
        }
    }
    return pos;
}

/* The lone trailing I of a capture with an odd number of floats is paired with whatever the
 * reference's static 64 KiB fread buffer holds one slot further (papr.c:35,101-103): the float the
 * previous full chunk left there, or 0.0 for a capture shorter than one chunk, overlaid with the
 * <4 trailing bytes of a ragged file (glibc fread stores the partial item). */
float papr_host_stale_q(const unsigned char *img, uint64_t bytes)
{
    const uint64_t chunk = 16384; /* CHUNK_SIZE, papr.c:30 */
    uint64_t nfloats = bytes / 4, last = nfloats - 1;
    uint64_t slot = last % chunk + 1, chunk_start = last - last % chunk;
    unsigned char q[4] = {0, 0, 0, 0};
    float f;
    if (chunk_start >= chunk) memcpy(q, img + 4 * (chunk_start - chunk + slot), 4);
    if (bytes % 4) memcpy(q, img + 4 * nfloats, bytes % 4);
    memcpy(&f, q, 4);
    return f;
}

/* papr.c:103-104 replayed literally on a short run of samples (the tiles of the exact sequential-sum
 * emulation that may cross a power of two): sum += (I*I)+(Q*Q), float products/sum, double accumulate. */
double papr_host_seq_add(double sum, const float *iq, uint64_t nsamples)
{
    for (uint64_t k = 0; k < nsamples; k++) {
        volatile float ii = iq[2 * k] * iq[2 * k];
        volatile float qq = iq[2 * k + 1] * iq[2 * k + 1];
        volatile float value = ii + qq;
        sum += value;
    }
    return sum;
}
