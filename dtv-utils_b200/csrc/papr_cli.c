/* papr — drop-in for the reference tool's executable: main() of papr.c:32 lives in libpapr_b200.so (papr_main).
 * Once the text is on stdout nothing is left to do, so the process leaves through _exit(): the CUDA
 * runtime's teardown at normal exit (~0.15 s of a 0.8 s run on a 16 GiB capture) frees nothing the
 * kernel does not reclaim anyway. */
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

#include "../../include/papr_b200.h"

int main(int argc, char **argv)
{
    /* CUDA initialisation time grows with the number of visible GPUs (seconds on an 8-GPU box): expose
     * only the ones this run uses, unless the user already restricted them.  This belongs to the
     * stand-alone executable, before CUDA is touched - the library entry point (papr_main) never edits
     * the environment of a host process that may have its own CUDA context. */
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        const char *v;
        char list[256];
        int ndev = (v = getenv("PAPR_B200_DEVICES")) ? atoi(v) : 1;
        int pos = 0, first = (v = getenv("PAPR_B200_DEVICE")) ? atoi(v) : 0;
        if (ndev < 1) ndev = 1;
        for (int d = 0; d < ndev && pos < 240; d++) pos += snprintf(list + pos, sizeof(list) - pos, d ? ",%d" : "%d", first + d);
        setenv("CUDA_VISIBLE_DEVICES", list, 1);
        setenv("PAPR_B200_DEVICE", "0", 1);
    }
    int rc = papr_main(argc, argv);
    fflush(stdout);
    fflush(stderr);
    _exit(rc);
}
