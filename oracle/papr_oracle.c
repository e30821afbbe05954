/*
 * oracle/papr_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see papr_oracle.h).
 *
 * Restates drmpeg/dtv-utils papr.c:32-196.  Written from the behaviour documented in SURVEY.md
 * §8a; every function cites the reference lines it follows.  Compile WITHOUT -march=native /
 * -ffast-math and with -ffp-contract=off: the reference's canonical x86-64 build rounds I*I,
 * Q*Q and their sum separately.
 */
#define _GNU_SOURCE
#include "papr_oracle.h"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

void papr_oracle_stats_init(papr_oracle_stats *st)
{
    memset(st, 0, sizeof(*st)); /* all-zero initial values, papr.c:37-49 */
}

/* papr.c:103 — three separately rounded float operations */
static inline float power_of(float i, float q)
{
    volatile float ii = i * i;
    volatile float qq = q * q;
    return ii + qq;
}

/* papr.c:102-128 */
void papr_oracle_pass1(const float *iq, int64_t nsamples, papr_oracle_stats *st)
{
    for (int64_t k = 0; k < nsamples; k++) {
        float re = iq[2 * k], im = iq[2 * k + 1];
        float value = power_of(re, im);
        st->sum += value;                         /* papr.c:104, file order */
        if (value > st->peak) {                   /* papr.c:105-108, strict: first wins */
            st->peak = value;
            st->peak_idx = st->n;
        }
        if (re > st->re_pos) { st->re_pos = re; st->re_pos_idx = st->n; } /* :110-113 */
        if (re < st->re_neg) { st->re_neg = re; st->re_neg_idx = st->n; } /* :114-117 */
        if (im > st->im_pos) { st->im_pos = im; st->im_pos_idx = st->n; } /* :119-122 */
        if (im < st->im_neg) { st->im_neg = im; st->im_neg_idx = st->n; } /* :123-126 */
        st->n++;                                  /* papr.c:127 */
    }
}

/* (int)x as the reference's x86-64 build evaluates it (cvttss2si: NaN / out of range -> INT_MIN) */
static int int_of_float_x86(float x)
{
    if (!(x >= -2147483648.0f && x < 2147483648.0f)) return (-2147483647 - 1);
    return (int)x;
}

int papr_oracle_levels(const papr_oracle_stats *st, int graph, double *avg_out, float *papr_out,
                       float *level, int cap)
{
    volatile double avg = st->sum / (double)st->n;            /* papr.c:131 / 164 */
    volatile double ratio = (double)st->peak / avg;
    float papr = (float)(10.0 * log10(ratio));                /* papr.c:134 / 165 */
    *avg_out = avg;
    *papr_out = papr;
    int top;                                                  /* loops run for i = 0..top */
    if (!graph) {
        top = int_of_float_x86(papr);                         /* papr.c:138 */
    } else {
        volatile float p10 = papr * 10.0f;                    /* papr.c:169, float multiply */
        top = int_of_float_x86(p10);
    }
    if (top < 0) return 0;
    int L = top + 1;
    if (!graph) {
        for (int i = 0; i < L && i < cap; i++) {
            volatile float x = (float)i / 10.0f;              /* papr.c:139 */
            level[i] = (float)(pow(10.0, (double)x) * avg);
        }
    } else {
        float index = 0.0f;                                   /* papr.c:168 */
        for (int i = 0; i < L; i++) {
            if (i < cap) {
                volatile float x = index / 10.0f;             /* papr.c:170 */
                level[i] = (float)(pow(10.0, (double)x) * avg);
            }
            index = (float)((double)index + 0.1);             /* papr.c:172, float drift */
        }
    }
    return L;
}

/* papr.c:147-151 (and 179-183) restated: c(v) = #{j : level[j] < v}; level_count[j] = #{v : c(v) > j} */
void papr_oracle_pass2(const float *iq, int64_t nsamples, const float *level, int L,
                       int64_t *level_count)
{
    if (L <= 0) return;
    int monotone = 1;
    for (int j = 1; j < L; j++)
        if (!(level[j] >= level[j - 1])) monotone = 0;
    if (!monotone) { /* the literal loop */
        for (int64_t k = 0; k < nsamples; k++) {
            float value = power_of(iq[2 * k], iq[2 * k + 1]);
            for (int j = 0; j < L; j++)
                if (value > level[j]) level_count[j]++;
        }
        return;
    }
    int64_t *hist = (int64_t *)calloc((size_t)L + 1, sizeof(int64_t));
    for (int64_t k = 0; k < nsamples; k++) {
        float value = power_of(iq[2 * k], iq[2 * k + 1]);
        int lo = 0, hi = L; /* first j with !(level[j] < value) */
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (level[mid] < value) lo = mid + 1; else hi = mid;
        }
        hist[lo]++;
    }
    int64_t above = 0;
    for (int j = L - 1; j >= 0; j--) {
        above += hist[j + 1];
        level_count[j] += above;
    }
    free(hist);
}

#define EMIT(...)                                                       \
    do {                                                                \
        int w_ = snprintf(out + pos, cap - (size_t)pos, __VA_ARGS__);   \
        if (w_ < 0 || (size_t)w_ >= cap - (size_t)pos) return -1;       \
        pos += w_;                                                      \
    } while (0)

long papr_oracle_format(const papr_oracle_stats *st, int graph, double avg, float papr,
                        const int64_t *level_count, int L, char *out, size_t cap)
{
    long pos = 0;
    if (cap == 0) return -1;
    out[0] = 0;
    if (!graph) {
        EMIT("Peak magnitude = %f\n", sqrt((double)st->peak));                     /* :132 */
        EMIT("average power = %lf, peak power = %f @ %lld\n\n", avg, (double)st->peak,
             (long long)(st->peak_idx * 8));                                        /* :133 */
        EMIT("Maximum PAPR = %f\n", (double)papr);                                  /* :135 */
        for (int i = 0; i < L; i++) {                                               /* :154-156 */
            volatile float frac = (float)level_count[i] / (float)st->n;
            EMIT("percentage above %d dB = %0.8f\n", i, (double)frac * 100.0);
        }
        EMIT("\n");                                                                 /* :157 */
        EMIT("peak real positive = %f, peak imaginary positive = %f\n", (double)st->re_pos,
             (double)st->im_pos);                                                   /* :158 */
        EMIT("peak real negative = %f, peak imaginary negative = %f\n\n", (double)st->re_neg,
             (double)st->im_neg);                                                   /* :159 */
        EMIT("peak real positive @ %lld, peak imaginary positive @ %lld\n",
             (long long)(st->re_pos_idx * 8), (long long)(st->im_pos_idx * 8 + 1)); /* :160 */
        EMIT("peak real negative @ %lld, peak imaginary negative @ %lld\n",
             (long long)(st->re_neg_idx * 8), (long long)(st->im_neg_idx * 8 + 1)); /* :161 */
    } else {
        for (int i = 0; i < L; i++) {                                               /* :187-190 */
            volatile float frac = (float)level_count[i] / (float)st->n;
            EMIT("%0.8f\n", (double)frac * 100.0);
        }
    }
    return pos;
}

/*
 * The reference reads with fread(buffer, 4, 16384, fp) into a static, zero-initialised buffer
 * (papr.c:35,101) and then walks i = 0,2,4.. < length touching buffer[i+1] (papr.c:102-103,119).
 * If the file holds an odd number of floats the last I is paired with whatever is at
 * buffer[length]: the float the previous full chunk left there, or 0.0 for a file shorter than one
 * chunk; glibc's fread also deposits the <4 trailing bytes of a ragged file there.
 */
static float stale_q_of(const unsigned char *img, size_t bytes)
{
    size_t nfloats = bytes / 4;
    size_t last = nfloats - 1;                   /* index of the lone I */
    size_t slot = last % PAPR_ORACLE_CHUNK + 1;  /* buffer index read as Q */
    unsigned char q[4] = {0, 0, 0, 0};
    size_t chunk_start = last - last % PAPR_ORACLE_CHUNK;
    if (chunk_start >= PAPR_ORACLE_CHUNK)
        memcpy(q, img + 4 * (chunk_start - PAPR_ORACLE_CHUNK + slot), 4);
    size_t ragged = bytes % 4;
    if (ragged) memcpy(q, img + 4 * nfloats, ragged);
    float f;
    memcpy(&f, q, 4);
    return f;
}

long papr_oracle_run_buffer(const void *file_image, size_t file_bytes, int graph, char *out,
                            size_t cap)
{
    const unsigned char *img = (const unsigned char *)file_image;
    size_t nfloats = file_bytes / 4;
    int64_t npairs = (int64_t)(nfloats / 2);
    const float *iq = (const float *)file_image;
    float tail[2] = {0, 0};
    int has_tail = (nfloats & 1) != 0;
    if (has_tail) {
        memcpy(&tail[0], img + 4 * (nfloats - 1), 4);
        tail[1] = stale_q_of(img, file_bytes);
    }
    papr_oracle_stats st;
    papr_oracle_stats_init(&st);
    papr_oracle_pass1(iq, npairs, &st);
    if (has_tail) papr_oracle_pass1(tail, 1, &st);

    double avg;
    float papr;
    int L = papr_oracle_levels(&st, graph, &avg, &papr, NULL, 0);
    float *level = NULL;
    int64_t *level_count = NULL;
    if (L > 0) {
        level = (float *)malloc((size_t)L * sizeof(float));
        level_count = (int64_t *)calloc((size_t)L, sizeof(int64_t));
        papr_oracle_levels(&st, graph, &avg, &papr, level, L);
        papr_oracle_pass2(iq, npairs, level, L, level_count);
        if (has_tail) papr_oracle_pass2(tail, 1, level, L, level_count);
    }
    long n = papr_oracle_format(&st, graph, avg, papr, level_count, L, out, cap);
    free(level);
    free(level_count);
    return n;
}

long papr_oracle_run_file(const char *path, int graph, char *out, size_t cap)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return -2;
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); return -2; }
    size_t bytes = (size_t)sb.st_size;
    void *img = NULL;
    if (bytes) {
        img = mmap(NULL, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
        if (img == MAP_FAILED) { close(fd); return -2; }
    }
    long n = papr_oracle_run_buffer(img, bytes, graph, out, cap);
    if (img) munmap(img, bytes);
    close(fd);
    return n;
}

/* SURVEY.md Appendix A */
static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline float siggen_component(uint64_t idx, unsigned c, uint64_t seed)
{
    int32_t s = -262140; /* 8 x 32767.5 */
    for (unsigned w = 0; w < 2; w++) {
        uint64_t z = mix64((4 * idx + 2 * c + w + 1) * 0x9E3779B97F4A7C15ULL + seed);
        s += (int32_t)((z & 0xffff) + ((z >> 16) & 0xffff) + ((z >> 32) & 0xffff) + (z >> 48));
    }
    return (float)s * 0x1p-19f; /* |s| < 2^24: exact */
}

void papr_oracle_siggen(float *iq, uint64_t first, uint64_t nsamples, uint64_t seed)
{
    for (uint64_t k = 0; k < nsamples; k++) {
        iq[2 * k] = siggen_component(first + k, 0, seed);
        iq[2 * k + 1] = siggen_component(first + k, 1, seed);
    }
}
