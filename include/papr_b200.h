/*
 * papr_b200.h — C ABI of the B200-native PAPR/CCDF engine (libpapr_b200.so).
 *
 * The reference tool (drmpeg/dtv-utils papr.c) has no library/plugin/FFI API: its only boundary
 * is the process, `int main(int argc, char **argv)` (papr.c:32) with the command line
 * `papr [-g] <infile>` (papr.c:53-98) and the text it prints (papr.c:132-135,154-161,186-190).
 * `papr_main` below IS that boundary.  The remaining entry points expose, step by step, the
 * stages that `main` inlines, so that a host in any language can bind them:
 * every declaration cites the reference lines it replaces.
 *
 * Conventions: plain pointers and sizes only; 0 = success, negative = error (text via
 * papr_last_error); one engine per GPU and per calling thread; "device" pointers are CUDA device
 * addresses on the engine's GPU, 16-byte aligned; a "sample" is one float32 I/Q pair (8 bytes,
 * the gr_complex file format, papr.c:101-103).  Sample indices are 0-based positions in the
 * capture; the reference prints index*8 as a byte offset (papr.c:133,160-161).
 */
#ifndef PAPR_B200_H
#define PAPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAPR_B200_ABI_VERSION 1
/* sum >= peak, so peak/avg <= N < 2^64 and PAPR < 192.7 dB: at most 1927 levels in -g mode */
#define PAPR_MAX_LEVELS 2048

enum {
    PAPR_OK = 0,
    PAPR_ERR_CUDA = -1,     /* CUDA runtime / driver failure (no GPU, OOM, launch error) */
    PAPR_ERR_ARG = -2,      /* bad argument (alignment, size, NULL) */
    PAPR_ERR_IO = -3,       /* cannot open / map the capture */
    PAPR_ERR_INTERNAL = -4, /* self-check failed (device/host level tables disagree) */
};

/* How the CCDF pass is scheduled against the statistics pass. */
enum {
    PAPR_MODE_AUTO = 0,    /* fused for large inputs, two-pass for small ones */
    PAPR_MODE_TWO_PASS = 1,/* stats pass, then histogram pass with the exact thresholds (16 B/sample) */
    PAPR_MODE_FUSED = 2,   /* one pass: thresholds predicted from a subsample, verified a posteriori,
                              automatic two-pass redo on a miss (~8.3 B/sample) */
};

/* State of the reference's first loop (papr.c:36-49,100-129) for one contiguous range of the
 * capture.  Ranges merge associatively in index order (papr_stats_merge). */
typedef struct papr_stats {
    uint64_t n;           /* offset: samples accumulated                    papr.c:37,127 */
    double   sum;         /* sum of float32 powers, as a double             papr.c:39,104 */
    float    peak;        /* largest I*I+Q*Q, 0.0 if none > 0               papr.c:40,105-108 */
    float    re_pos;      /* peak_real_pos                                  papr.c:42,110-113 */
    float    im_pos;      /* peak_imag_pos                                  papr.c:43,119-122 */
    float    re_neg;      /* peak_real_neg                                  papr.c:44,114-117 */
    float    im_neg;      /* peak_imag_neg                                  papr.c:45,123-126 */
    uint32_t flags;       /* PAPR_FLAG_* */
    uint64_t peak_idx;    /* first sample index attaining each extreme (0 if none) papr.c:38,46-49 */
    uint64_t re_pos_idx, im_pos_idx, re_neg_idx, im_neg_idx;
} papr_stats;

#define PAPR_FLAG_NONFINITE 1u /* sum is NaN/Inf: no CCDF lines are printed (papr.c:136-141 with NaN) */

/* Everything the reference prints. */
typedef struct papr_result {
    papr_stats stats;
    double   avg;                          /* sum / offset                  papr.c:131,164 */
    float    papr;                         /* 10*log10(peak/avg) as float   papr.c:134,165 */
    int32_t  graph;                        /* -g given                      papr.c:75-78 */
    int32_t  nlevels;                      /* (int)papr+1 or (int)(papr*10)+1, 0 if the loops do not run */
    float    level[PAPR_MAX_LEVELS];       /* power thresholds              papr.c:139,170 */
    int64_t  level_count[PAPR_MAX_LEVELS]; /* samples with power > level[j] papr.c:147-151,179-183 */
    /* diagnostics (not printed) */
    int32_t  mode_used;                    /* PAPR_MODE_TWO_PASS or PAPR_MODE_FUSED */
    int32_t  fused_miss;                   /* 1 if the fused pass missed and the exact pass re-ran; 2 if only the level counts were taken again */
    float    device_ms;                    /* GPU time of the analysis (CUDA events on the engine stream) */
    float    scan_ms;                      /* GPU time of the dominant scan kernel launch(es) */
    uint32_t kernel_launches;              /* kernels launched for this analysis */
    uint32_t sum_path;                     /* how stats.sum was obtained: 0 = any fixed order (exact_sum off, or NaN/Inf
                                              powers: nothing to emulate); 1 = papr.c:104 emulated inside the fused sweep and
                                              chained on the device; 2 = two-sweep emulation (bits 8-15: why the device chain
                                              declined, 0 if it was not tried); 3 = emulated chunk by chunk while streaming in */
    uint64_t h2d_bytes, d2h_bytes;         /* bytes moved across PCIe by this call */
} papr_result;

typedef struct papr_engine papr_engine;

/* ---- lifetime ------------------------------------------------------------------------------ */
int  papr_abi_version(void);
/* device < 0 : current device.  Allocates pinned staging and device work buffers lazily. */
int  papr_engine_create(int device, papr_engine **out);
void papr_engine_destroy(papr_engine *e);
const char *papr_last_error(const papr_engine *e); /* e may be NULL: last create() error */
/* cudaStream_t the engine launches on (for callers that time or order against it). */
void *papr_engine_stream(papr_engine *e);
/* Tunables by name ("mode", "presample_stride", "window_sigmas", "chunk_bytes", "staging_threads",
 * "fused_min_samples", "fine_bytes_log2", "predict_bias" (test hook: scales the fused mode's predicted
 * mean; anything but 1.0 provokes the a-posteriori miss and the exact redo), "max_resident_bytes": host-side captures larger than this
 * (default 0 = what the GPU has free, less 1 GiB) are not kept resident in HBM but streamed twice,
 * like the reference reads its file twice (papr.c:142); "exact_sum": != 0 (default) = emulate the reference's sequential
 * double sum (papr.c:104) bit for bit on every path - inside the fused sweep for large device-resident shards
 * (papr_result.sum_path says which way), 0 = any fixed summation order; "o_direct": 1 = regular files are also opened with O_DIRECT and every
 * 4096-byte-aligned piece is read straight into the pinned staging slots, bypassing the page cache (silently buffered
 * where the file system refuses O_DIRECT, e.g. tmpfs); "p2p_chain": whether papr_shard_analyze_p2p chains the
 * sequential sum on the device: -1 (default) = if this rank's shard qualifies (>= 8192 samples); 0 = never; 1 = fail if
 * it does not qualify.  A rank that does not chain still takes part in the chain exchange and declines there, so every
 * rank reports the same fall-back (sum_path 2) and the ranks need not agree on anything beforehand;
 * "epilogue_bias": test hook (1.0), scales the fixed-order sum from which the levels are derived while the chain is still
 * running (a disagreement with the levels of the chained sum must be caught: fused_miss 2); "xchg_timeout_s": how long the in-kernel
 * peer exchange waits for a rank before ALL ranks give up (default 30)).  Returns PAPR_ERR_ARG for an unknown name. */
int  papr_engine_set(papr_engine *e, const char *name, double value);

/* ---- the process boundary: replaces main(), papr.c:32-196 ------------------------------------ */
/* Same argv surface, same stdout/stderr text, returns the exit status instead of calling exit():
 * 0 on success, 255 (the reference's exit(-1)) for usage / open errors, 1 for GPU failures.
 * Extra knobs come from the environment only (PAPR_B200_DEVICES, PAPR_B200_MODE, ...), because
 * any extra argv makes the reference print its usage text. */
int papr_main(int argc, char **argv);

/* ---- whole analysis: replaces papr.c:100-190 minus the printf calls --------------------------- */
/* Host buffer holding the FILE IMAGE (any length; a trailing lone I is paired with the stale Q
 * exactly as the reference's chunked fread does, papr.c:100-103).  Pinned or pageable; streamed to
 * the GPU in chunks (cudaMemcpyAsync straight from pinned memory, through a pinned staging ring
 * otherwise) with the statistics pass and the sequential-sum emulation running on each chunk as it
 * lands, so that only the CCDF pass over the now-resident shard remains after the last byte. */
int papr_analyze_host(papr_engine *e, const void *file_image, uint64_t file_bytes, int graph,
                      papr_result *out);
/* Device-resident complete samples. */
int papr_analyze_device(papr_engine *e, const float *d_iq, uint64_t nsamples, int graph,
                        papr_result *out);
/* The same for a file: regular files are read with pread() straight into the pinned staging ring
 * (no mapping).  FIFOs, pipes and other non-seekable inputs - which the reference cannot read at all, it
 * rewinds its file (papr.c:142) - are read ONCE: each piece goes from read(2) through a pinned slot to the
 * GPU while pass 1 runs on the chunks already there, and the chunks stay resident for the CCDF pass (the
 * stream must fit the GPU's free memory).  papr_analyze_fd takes an open descriptor (not closed). */
int papr_analyze_file(papr_engine *e, const char *path, int graph, papr_result *out);
int papr_analyze_fd(papr_engine *e, int fd, int graph, papr_result *out);

/* ---- one capture over several GPUs of this box (single process) --------------------------------- */
/* Byte-range shards (whole 64 MiB chunks), one engine and one host thread per GPU; pass-1 states
 * folded in range order, the sequential sum chained across shards, one reduction of the level counts:
 * on the host by default (they are 16 KB that every shard copies back anyway), or with one
 * ncclAllReduce over NVLink when PAPR_B200_EXCHANGE=nccl (libnccl.so.2 is dlopen()ed; communicator
 * setup costs tens of seconds on this pool).  A device may be listed twice
 * (virtual shards).  devices == NULL means 0..ndev-1. */
typedef struct papr_multi papr_multi;
int  papr_multi_create(int ndev, const int *devices, papr_multi **out);
void papr_multi_destroy(papr_multi *m);
int  papr_multi_set(papr_multi *m, const char *name, double value);       /* papr_engine_set on every shard */
int  papr_multi_analyze_host(papr_multi *m, const void *file_image, uint64_t file_bytes, int graph,
                             papr_result *out);
int  papr_multi_analyze_file(papr_multi *m, const char *path, int graph, papr_result *out);
int  papr_multi_analyze_fd(papr_multi *m, int fd, int graph, papr_result *out); /* a pipe is read by the first GPU */
const char *papr_multi_last_error(const papr_multi *m);
const char *papr_multi_exchange(const papr_multi *m); /* "nccl" or why the host summed the counts */

/* ---- the stages, for sharded (multi-GPU / multi-process) callers ------------------------------ */
/* papr.c:100-129 over [first_index, first_index+nsamples) held at d_iq.  Synchronous. */
int papr_stats_device(papr_engine *e, const float *d_iq, uint64_t nsamples, uint64_t first_index,
                      papr_stats *out);
/* Same for a shard held in HOST memory (file-image semantics as papr_analyze_host): streams it to the
 * GPU with the statistics pass overlapped and leaves it resident; *d_iq / *nsamples describe the
 * resident copy (engine-owned, valid until the next host-path call) for papr_ccdf_device. */
int papr_stats_host(papr_engine *e, const void *file_image, uint64_t file_bytes, uint64_t first_index,
                    papr_stats *out, const float **d_iq, uint64_t *nsamples);
/* Fold `next` (the range immediately after `acc`) into `acc`; first occurrence wins ties. */
void papr_stats_merge(papr_stats *acc, const papr_stats *next);
/* papr.c:131,134,136-141 / 164-173: avg, papr and the threshold table from merged stats, with the
 * host libm exactly as the reference evaluates them.  Returns the number of levels. */
int papr_levels(const papr_stats *st, int graph, double *avg, float *papr, float *level, int cap);
/* papr.c:143-153 / 175-185: level_count[j] += #{samples of this range with power > level[j]}. */
int papr_ccdf_device(papr_engine *e, const float *d_iq, uint64_t nsamples, const float *level,
                     int nlevels, int64_t *level_count);
/* Fused single-pass variant for a shard: phase 1 estimates (sum, sum of squares, blocks) of block
 * sums from a strided subsample; the caller sums pre[4] over ranks; phase 2 scans the shard once;
 * phase 3, given the merged stats of ALL ranks, produces this shard's counts or reports a miss
 * (returns 1) in which case the caller falls back to papr_ccdf_device on every rank. */
int papr_fused_presample(papr_engine *e, const float *d_iq, uint64_t nsamples, int graph, double pre[4]);
int papr_fused_scan(papr_engine *e, const float *d_iq, uint64_t nsamples, uint64_t first_index,
                    const double pre[4], int graph, papr_stats *out);
int papr_fused_counts(papr_engine *e, const papr_stats *merged, int graph, int64_t *level_count);

/* Stream-ordered variants of the shard stages: nothing synchronises with the host until
 * papr_shard_finish, so a caller (one process per GPU) interleaves its collectives on
 * papr_engine_stream() directly on the engine's device buffers:
 *     presample_async -> all-reduce(sum) of PAPR_BUF_PRESAMPLE (4 doubles)          [fused only]
 *     scan_async      -> all-gather of PAPR_BUF_LOCAL_STATS (opaque, device format) into d_all_stats
 *     counts_async    -> all-reduce(sum) of PAPR_BUF_COUNTS (PAPR_MAX_LEVELS+1 u64: counts + miss word)
 *     finish          -> result of the WHOLE capture; returns 1 if any rank missed (then call
 *                        counts_async with fused = 0, reduce and finish again). */
enum { PAPR_BUF_PRESAMPLE = 0, PAPR_BUF_LOCAL_STATS = 1, PAPR_BUF_COUNTS = 2 };
int papr_engine_device_buffer(papr_engine *e, int which, void **ptr, uint64_t *bytes);
int papr_shard_presample_async(papr_engine *e, const float *d_iq, uint64_t nsamples, int graph);
int papr_shard_scan_async(papr_engine *e, const float *d_iq, uint64_t nsamples, uint64_t first_index, int graph,
                          int fused);
int papr_shard_counts_async(papr_engine *e, const void *d_all_stats, int nparts, int graph, int fused,
                            const float *d_iq, uint64_t nsamples);
int papr_shard_finish(papr_engine *e, int graph, papr_result *out);

/* The same sharded analysis with the four exchanges (presample, pass-1 states, the chain of the sequential
 * sum, level counts) fused INTO the kernels over peer memory (NVLink,
 * one process per GPU): every engine owns a small window in device memory that the other ranks map
 * with cudaIpc; the producing kernel stores its values into every peer's window and raises a flag,
 * the consuming kernel polls its own window - no collective launches in between.
 *   export  -> 64-byte cudaIpcMemHandle_t of this engine's window; the caller gathers all ranks'
 *   attach  -> maps the other ranks' windows (world <= 16; every rank must attach before any analyses)
 *   papr_shard_analyze_p2p -> called by EVERY rank with its byte range; the whole-capture result on
 *              every rank.  Fused mode; on a miss every rank re-runs the exact pass (same decision
 *              everywhere, because the merged statistics and the summed status word are identical).
 *              Shards may differ in size; a rank whose shard cannot chain the sequential sum on the device
 *              declines inside the exchange and every rank reports sum_path 2.
 * A rank that never calls leaves ALL ranks (itself included, should it arrive late) with PAPR_ERR_INTERNAL
 * after the device-side timeout (tunable "xchg_timeout_s", default 30 s). */
int papr_xchg_export(papr_engine *e, void *handle64);
int papr_xchg_attach(papr_engine *e, int rank, int world, const void *handles /* world x 64 bytes */);
/* unmaps the peers' windows; detach on every rank, synchronise, THEN destroy the engines */
int papr_xchg_detach(papr_engine *e);
int papr_shard_analyze_p2p(papr_engine *e, const float *d_iq, uint64_t nsamples, uint64_t first_index, int graph,
                           papr_result *out);

/* papr.c:104 for a capture sharded over ranks: the reference's SEQUENTIAL double sum, bit for bit
 * (DESIGN.md section 5).  On the resident shard [d_iq, d_iq + 2*nsamples) of every rank:
 *   prepare -> *approx = sum of this shard (any order); the caller all-gathers these and forms
 *              pre = the sum of approx over the LOWER ranks;
 *   runs    -> places the shard's 32768-sample tiles in binades of the running sum (from pre) and
 *              reduces each tile to its parity-transducer run on the GPU; returns 1 if not applicable
 *              (a non-finite tile sum: the reference's sum is NaN/Inf as well);
 *   chain   -> rank after rank in index order: *state is the exact running sum after all lower
 *              ranks (0.0 on rank 0) on entry and after this shard on return; the caller hands it on.
 * The last rank's state replaces papr_stats.sum of the merged statistics. */
int papr_seqsum_prepare(papr_engine *e, const float *d_iq, uint64_t nsamples, double *approx);
int papr_seqsum_runs(papr_engine *e, const float *d_iq, uint64_t nsamples, double pre);
int papr_seqsum_chain(papr_engine *e, const float *d_iq, uint64_t nsamples, double *state);

/* papr.c:132-135,154-161 / 186-190: the exact stdout text.  Returns its length, or <0. */
long papr_format(const papr_result *r, char *out, size_t cap);
/* Fill avg/papr/levels of `r` from r->stats (papr_levels) — for callers that merged shards. */
int papr_result_finish(papr_result *r, int graph);

/* ---- PAPR reduction by tone reservation --------------------------------- */
/* What `dtv.dvbt2_paprtr_cc(..., vclip, iterations, fftsize)` of the reference's DVB-T2 chain does
 * (dvbt2-blade.py:52-54,129; the block is GNU Radio gr-dtv's, the algorithm EN 302 755 clause 9.6.2.1), on
 * nsym time-domain symbols of fft_size complex samples held at d_symbols (corrected in place):
 * d_kernel = the reference kernel p (fft_size complex: IFFT of 1 on every reserved tone, p[0] = 1),
 * d_tones = the reserved carrier indices (FFT bin numbers), amax = bound on what one reserved tone may
 * carry.  d_tone_values (nsym x ntones complex) receives the reserved tones' values in the kernel's
 * normalisation, d_iterations (nsym ints, may be NULL) the iterations each symbol took. */
int papr_tr_reduce_device(papr_engine *e, float *d_symbols, int nsym, int fft_size, const float *d_kernel,
                          const int *d_tones, int ntones, float vclip, int iterations, float amax,
                          float *d_tone_values, int *d_iterations);

/* ---- synthetic captures (SURVEY.md Appendix A generator, bit-identical to its C/numpy twins) -- */
int papr_siggen_device(papr_engine *e, float *d_iq, uint64_t first_index, uint64_t nsamples,
                       uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif /* PAPR_B200_H */
