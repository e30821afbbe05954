/*
 * oracle/papr_oracle_main.c — TEST INFRASTRUCTURE (see papr_oracle.h).
 *   papr_oracle [-g] <infile>                 same surface as the reference tool (papr.c:53-98)
 *   papr_oracle --gen <seed> <first> <nsamples> <outfile>   Appendix-A synthetic capture
 */
#include "papr_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv)
{
    if (argc == 6 && strcmp(argv[1], "--gen") == 0) {
        uint64_t seed = strtoull(argv[2], NULL, 0), first = strtoull(argv[3], NULL, 0);
        uint64_t n = strtoull(argv[4], NULL, 0);
        FILE *fp = fopen(argv[5], "wb");
        if (!fp) { perror(argv[5]); return 1; }
        enum { B = 1 << 16 };
        float *buf = (float *)malloc(sizeof(float) * 2 * B);
        for (uint64_t k = 0; k < n; k += B) {
            uint64_t m = n - k < B ? n - k : B;
            papr_oracle_siggen(buf, first + k, m, seed);
            fwrite(buf, 8, m, fp);
        }
        free(buf);
        fclose(fp);
        return 0;
    }
    int graph = 0;
    const char *path;
    if (argc == 2) {
        path = argv[1];
    } else if (argc == 3 && argv[1][0] == '-') {
        for (size_t i = 1; i < strlen(argv[1]); i++) {
            if (argv[1][i] == 'g' || argv[1][i] == 'G') graph = 1;
            else fprintf(stderr, "Unsupported Option: %c\n", argv[1][i]);
        }
        path = argv[2];
    } else {
        fprintf(stderr, "usage: papr -g <infile>\nOptions:\n\tg = graph suitable output\n");
        exit(-1);
    }
    size_t cap = 1 << 20;
    char *out = (char *)malloc(cap);
    long n = papr_oracle_run_file(path, graph, out, cap);
    if (n == -2) {
        fprintf(stderr, "Cannot open bitstream file <%s>\n", path);
        exit(-1);
    }
    if (n < 0) return 2;
    fwrite(out, 1, (size_t)n, stdout);
    free(out);
    return 0;
}
