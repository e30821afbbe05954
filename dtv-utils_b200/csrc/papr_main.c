/*
 * papr_main.c — the process boundary of the reference tool (drmpeg/dtv-utils papr.c:32-98,192-195):
 * same argv surface, same stderr texts, same exit statuses; the number crunching goes to the GPU
 * engine and the text comes from papr_format().  There is deliberately NO CPU fallback: without a
 * usable B200 the tool fails loudly.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include <fcntl.h>
#include <unistd.h>

#include "../../include/papr_b200.h"

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

/* papr.c:132-135,154-161 / 186-190: the text goes out as soon as the numbers exist, before the GPU
 * resources (a resident copy of the capture among them) are torn down */
static void print_result(const papr_result *r)
{
    size_t cap = 256 + (size_t)r->nlevels * 64 + 1024;
    char *text = (char *)malloc(cap);
    long len = text ? papr_format(r, text, cap) : -1;
    if (len > 0) fwrite(text, 1, (size_t)len, stdout);
    fflush(stdout);
    free(text);
}

static void usage(void)
{
    fprintf(stderr, "usage: papr -g <infile>\n"); /* papr.c:54-56, 86-88 */
    fprintf(stderr, "Options:\n");
    fprintf(stderr, "\tg = graph suitable output\n");
}

int papr_main(int argc, char **argv)
{
    int graph = 0;
    const char *path;
    if (argc != 2 && argc != 3) { /* papr.c:53-58 */
        usage();
        return 255; /* exit(-1) */
    }
    if (argc == 2) {
        path = argv[1]; /* papr.c:59-67 (even if it is literally "-g") */
    } else {
        if (argv[1][0] != '-') { /* papr.c:84-91 */
            usage();
            return 255;
        }
        for (size_t i = 1; i < strlen(argv[1]); i++) { /* papr.c:70-83 */
            if (argv[1][i] == 'g' || argv[1][i] == 'G')
                graph = 1;
            else
                fprintf(stderr, "Unsupported Option: %c\n", argv[1][i]); /* warn and continue */
        }
        path = argv[2];
    }
    /* papr.c:62-66 / 93-97: same check, same message.  Opened ONCE and handed down as a descriptor: a named
     * FIFO must not be opened and closed just to probe it (its writer would see the reader disappear) */
    const int fd = open(path, O_RDONLY);
    if (fd < 0) {
        fprintf(stderr, "Cannot open bitstream file <%s>\n", path);
        return 255;
    }

    const char *v;
    papr_result *r = (papr_result *)calloc(1, sizeof(*r));
    int ndev = (v = getenv("PAPR_B200_DEVICES")) ? atoi(v) : 1; /* GPUs to shard the capture over */
    if (ndev < 1) ndev = 1;
    if (ndev > 1) {
        papr_multi *m = NULL;
        /* stdout belongs to the reference's text: keep any NCCL banner / debug output off it */
        setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
        if ((v = getenv("NCCL_DEBUG")) && strcasecmp(v, "VERSION") == 0)
            setenv("NCCL_DEBUG", "WARN", 1); /* the VERSION level ignores NCCL_DEBUG_FILE and prints on stdout */
        int devs[64], first = (v = getenv("PAPR_B200_DEVICE")) ? atoi(v) : 0; /* first GPU of the run */
        if (ndev > 64) ndev = 64;
        for (int d = 0; d < ndev; d++) devs[d] = first + d;
        if (papr_multi_create(ndev, devs, &m) != PAPR_OK) {
            fprintf(stderr, "papr: GPU engine unavailable: %s\n", papr_multi_last_error(NULL));
            free(r);
            close(fd);
            return 1;
        }
        if ((v = getenv("PAPR_B200_CHUNK_MB"))) papr_multi_set(m, "chunk_bytes", atof(v) * 1048576.0);
        if ((v = getenv("PAPR_B200_STAGING_THREADS"))) papr_multi_set(m, "staging_threads", atof(v));
        if ((v = getenv("PAPR_B200_EXACT_SUM"))) papr_multi_set(m, "exact_sum", atof(v));
        if ((v = getenv("PAPR_B200_ODIRECT"))) papr_multi_set(m, "o_direct", atof(v));
        int rc = papr_multi_analyze_fd(m, fd, graph, r);
        if (rc != PAPR_OK) {
            fprintf(stderr, "papr: %s\n", papr_multi_last_error(m));
            papr_multi_destroy(m);
            free(r);
            close(fd);
            return 1;
        }
        if (getenv("PAPR_B200_STATS"))
            fprintf(stderr, "papr_b200: %d shards, exchange=%s, wall_ms=%.3f launches=%u h2d=%llu\n", ndev,
                    papr_multi_exchange(m), r->device_ms, r->kernel_launches, (unsigned long long)r->h2d_bytes);
        print_result(r);
        papr_multi_destroy(m);
    } else {
        papr_engine *e = NULL;
        const char *dev = getenv("PAPR_B200_DEVICE");
        const double t0 = now_ms();
        if (papr_engine_create(dev ? atoi(dev) : -1, &e) != PAPR_OK) {
            fprintf(stderr, "papr: GPU engine unavailable: %s\n", papr_last_error(NULL));
            free(r);
            close(fd);
            return 1;
        }
        if ((v = getenv("PAPR_B200_CHUNK_MB"))) papr_engine_set(e, "chunk_bytes", atof(v) * 1048576.0);
        if ((v = getenv("PAPR_B200_STAGING_THREADS"))) papr_engine_set(e, "staging_threads", atof(v));
        if ((v = getenv("PAPR_B200_EXACT_SUM"))) papr_engine_set(e, "exact_sum", atof(v)); /* default: on for files */
        if ((v = getenv("PAPR_B200_ODIRECT"))) papr_engine_set(e, "o_direct", atof(v)); /* NVMe -> pinned staging, no page cache */
        if ((v = getenv("PAPR_B200_MAX_RESIDENT_MB"))) papr_engine_set(e, "max_resident_bytes", atof(v) * 1048576.0);
        const double t1 = now_ms();
        int rc = papr_analyze_fd(e, fd, graph, r);
        const double t2 = now_ms();
        if (rc != PAPR_OK) {
            fprintf(stderr, "papr: %s\n", papr_last_error(e));
            papr_engine_destroy(e);
            free(r);
            close(fd);
            return 1;
        }
        print_result(r);
        const double t3 = now_ms();
        papr_engine_destroy(e);
        if (getenv("PAPR_B200_STATS"))
            fprintf(stderr, "papr_b200: create_ms=%.1f analyze_ms=%.1f destroy_ms=%.1f device_ms=%.3f scan_ms=%.3f launches=%u h2d=%llu\n",
                    t1 - t0, t2 - t1, now_ms() - t3, r->device_ms, r->scan_ms, r->kernel_launches,
                    (unsigned long long)r->h2d_bytes);
    }
    free(r);
    close(fd);
    return 0;
}
