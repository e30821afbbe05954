// papr_device.cuh — device-side data layout shared by the kernels and the engine.
//
// HBM layout (per GPU):
//   capture shard      float32 I/Q interleaved, exactly the file bytes (gr_complex, papr.c:101-103)
//   CTA partials       PaprCtaPartial[grid]        running pass-1 state of each resident CTA
//   cell histogram     u64[16384]                 samples per "cell" = float32 bit pattern >> sh
//   ambiguity map      u32[16384]                 cell -> 1 + first fine-table counter, 0 = unambiguous
//   fine table         u64[slots << sh]           one counter per float32 value of an ambiguous cell
//   tables             double pow10[L], ratio_min[L] per mode (host libm, built once)
#pragma once
#include <stdint.h>
#include "../../include/papr_b200.h"

#ifndef PAPR_THREADS
#define PAPR_THREADS 1024                      // threads per CTA of the scan kernels
#endif
#define PAPR_WARPS (PAPR_THREADS / 32)
#ifndef PAPR_U
#define PAPR_U 4                               // float4 loads in flight per lane per batch
#endif
#ifndef PAPR_CTAS_PER_SM
#define PAPR_CTAS_PER_SM 1
#endif
#define PAPR_BATCH_VEC (32 * PAPR_U)           // float4 per warp batch
#define PAPR_BATCH_SAMPLES (2 * PAPR_BATCH_VEC)// 256 samples = 2 KiB contiguous per warp batch
#define PAPR_NCELLS_MAX 16384                  // cells of the shared-memory histogram
#define PAPR_SH_MIN 12                         // finest cell: 2^12 float32 values (2048 cells / octave)
#define PAPR_SH_MAX 23
#define PAPR_NTRACK 5                          // peak power, +I, -I, +Q, -Q
#define PAPR_SEQ_TILE 32768                    // samples per tile of the exact sequential-sum emulation
#define PAPR_SEQ_ZERO 30000                    // tile code: all samples zero (identity)
#define PAPR_SEQ_DIRTY 30001                   // tile code: may cross a power of two -> replayed on the host

// order of the five extreme trackers everywhere on the device
enum { TR_PEAK = 0, TR_RE_POS = 1, TR_RE_NEG = 2, TR_IM_POS = 3, TR_IM_NEG = 4 };

// pass-1 state of one shard (or the merge of several); papr.c:36-49
struct PaprDevStats {
    double sum;
    unsigned long long n;
    unsigned long long idx[PAPR_NTRACK]; // global sample index of the first occurrence
    int val[PAPR_NTRACK];                // float32 bits of the magnitude (>= 0; 0 = reference's 0.0 init)
    unsigned int flags;
};

// running pass-1 state of one CTA, persistent across the chunk launches of a streamed shard
// (all-zero bytes = the reference's initial values, papr.c:37-49, so a memset resets it)
struct PaprCtaPartial {
    double sum;
    unsigned long long idx[PAPR_NTRACK];
    int val[PAPR_NTRACK];
    unsigned int pad;
};

enum { PLAN_HIST = 1, PLAN_BSEARCH = 2 };

struct PaprPlan {
    int sh;             // cell = (float bits >> sh) - cell_base
    int cell_base;
    int ncells;         // <= PAPR_NCELLS_MAX
    int n_amb;          // ambiguous cells that own a slot in the fine table
    int status;         // PLAN_HIST: cell histogram valid; PLAN_BSEARCH: generic kernel must run instead
    int levels_covered; // fused mode: levels whose window fits the cell range / fine table
    float window;       // fused mode: relative half-width of the threshold windows
    int pad;            // peer-memory exchange: 1 = a peer's publication did not arrive in time
    double avg_pred;    // fused mode: predicted mean power
    double wcv;         // fused mode: sigmas x relative std of one warp batch's power sum (how far the sum
                        // of the first g samples may stray from g x avg_pred: wcv / sqrt(g / batch))
};

struct PaprDevLevels {
    double avg;
    double ratio;       // (double)peak / avg
    int L;
    int graph;
    int pad[2];
    float level[PAPR_MAX_LEVELS];
};

// status word written by the resolve kernel; it lives right after the level counts
// (counts[PAPR_MAX_LEVELS]) so that one all-reduce(sum) carries both across ranks
enum { RES_MISS = 1 };

struct PaprScanArgs {
    const float *iq;               // 16-byte aligned
    unsigned long long nsamples;   // complete samples in this launch, < 2^32
    unsigned long long first_index;// global sample index of iq[0]
    PaprCtaPartial *wp;            // [gridDim.x]
    const PaprPlan *plan;
    const unsigned *fine_base;     // [PAPR_NCELLS_MAX] 1 + first fine-table counter of the cell, 0 = none
    unsigned long long *g_hist;    // [PAPR_NCELLS_MAX]
    unsigned long long *g_fine;    // [n_amb << sh]
    unsigned long long *g_over;    // samples above the cell range
};

// ---- the reference's SEQUENTIAL double sum (papr.c:104) inside the fused scan (papr_exact.cu) -------
// The TMA-fed scan gives every lane 16 consecutive samples per warp batch and every warp whole "tiles"
// of 8 batches; per tile it emits a RUN - the increment of the running sum for an even and for an odd
// entry state, valid inside one binade [2^k, 2^(k+1)) of the running sum (see papr_seqsum_kernel).  k is
// predicted per tile from the presample mean; tiles that may straddle a power of two carry one run per
// candidate binade and per batch ("multi" tiles).  A single-CTA kernel then chains everything in file
// order, resolving each binade crossing down to the 16 samples in which it happens.
#define XT_RUN 16                                              // consecutive samples per lane and batch (one 128-byte row)
#define XT_BATCH_SAMPLES (32 * XT_RUN)                         // 512 samples = 4 KiB = one TMA box of 32 rows
#define XT_TILE_BATCHES 8
#define XT_TILE_SAMPLES (XT_TILE_BATCHES * XT_BATCH_SAMPLES)   // 4096 samples = 32 KiB per warp tile
#define XT_SUPER_TILES 32                                      // tiles per super-tile record (one lane each in the compose kernel)
#define XT_SUPER_SAMPLES (XT_SUPER_TILES * XT_TILE_SAMPLES)
#define XT_HYPER_SUPERS 32                                     // super-tiles per hyper-tile record (4M samples)
#define XT_HYPER_SAMPLES ((unsigned long long)XT_HYPER_SUPERS * XT_SUPER_SAMPLES)
#define XT_NCELLS PAPR_NCELLS_MAX                              // cells of the TMA-fed sweep's histogram
#define XT_MAX_CAND 4                                          // candidate binades of a multi tile
#define XT_KBIAS 1100                                          // tile code: k + XT_KBIAS in bits 0-11 ...
#define XT_CODE_NC(c) (((c) >> 12) & 7)                        // ... candidates in bits 12-14 (0 = none) ...
#define XT_CODE_LITERAL 0x8000                                 // ... bit 15: summed literally, e0 = exit state ...
#define XT_CODE_SLOT(c) ((unsigned)(c) >> 16)                  // ... bits 16-31: slot in the multi log
#define XT_CODE_K(c) (((c) & 0xfff) - XT_KBIAS)
#define XT_MAX_CROSS 44                                        // super-tiles per shard in which the running sum changes binade

struct PaprTileRun { double e0, e1; };                         // increments for an even / odd entry state

#define XT_SUPER_CAND 3
struct PaprSuperRec {                                          // 32 tiles composed (papr_xt_compose_kernel)
    double asum;                                               // approximate sum of the super-tile's powers
    int k_lo, nc;                                              // binades every tile of it has a run for: k_lo .. k_lo+nc-1
    PaprTileRun r[XT_SUPER_CAND];                              // ... and the composed run for each of them
};

struct PaprExactArgs {
    unsigned long long g_first;    // whole-capture index of this launch's sample 0 (prediction of the running sum)
    unsigned tile_base;            // index of this launch's first tile in the shard's tile arrays
    int literal_tile0;             // 1: this launch starts the capture (running sum exactly 0): tile 0 is summed literally
    PaprTileRun *tile_run;         // [ntiles]
    int *tile_code;                // [ntiles]
    PaprTileRun *multi;            // [multi_cap][XT_MAX_CAND][XT_TILE_BATCHES] per-batch runs of the multi tiles
    PaprTileRun *multi_tile;       // [multi_cap][XT_MAX_CAND] the same composed over the tile
    unsigned *multi_count;
    unsigned multi_cap;
};

// what the chain of one shard boils down to: a short list of items applied in order to the running sum
enum { XT_IT_SEG = 1, XT_IT_LIT = 2, XT_IT_ABS = 3 };
struct PaprChainItem {
    int type;       // SEG: a run (e0, e1) valid in binade k; LIT: the one power e0 with which the running sum enters
    int k;          //      binade k, added literally; ABS: the running sum becomes e0 (valid for an entry state of 0)
    double e0, e1;
};
#define XT_MAX_ITEMS 96   // per shard, after runs of the same binade have been merged (a shard of 2^37 samples passes
                          // < 40 binades above its first tile: one LIT and one or two SEG items each)
#define XT_MAX_RAW 384    // ... before the merge
enum { XT_OK = 0, XT_FALLBACK = 1, XT_NONFINITE = 2 };
struct PaprChainList {
    int n;
    int status;     // XT_*
    int why;        // diagnostic: which check sent the chain to the fall-back
    int pad;
    double approx;  // approximate sum of the shard (what the tree sum used to be)
    double exact;   // (after the walk) the running sum after this shard
    PaprChainItem item[XT_MAX_ITEMS];
};

// ---- peer-memory exchange between the ranks of one box (NVLink, cudaIpc-mapped windows) -----------
// Every GPU owns one window; slot q of it is written by rank q (its stores travel over NVLink),
// flag[k][q] carries the sequence number of rank q's latest publication of kind k.  The three
// latency-sized exchanges of a sharded analysis happen INSIDE the kernels that produce / consume
// the values: publish to every peer's window, release-store the flag, acquire-poll the own window.
#define PAPR_XCHG_MAX_RANKS 16
enum { XK_PRE = 0, XK_STATS = 1, XK_COUNTS = 2, XK_CHAIN = 3, XK_KINDS = 4 };

struct PaprXchgSlot {
    double pre[4];                                   // presample {sum, sum of squares, count, -}
    PaprDevStats stats;                              // pass-1 state of the shard
    // level counts + status word; two buffers used alternately (publication seq & 1): an analysis that
    // publishes counts twice (fused miss -> exact redo) never overwrites what a peer may still be summing
    unsigned long long counts[2][PAPR_MAX_LEVELS + 1];
    PaprChainList chain;                             // what the shard's samples do to the running sum (papr_exact.cu)
};

struct PaprXchg {
    unsigned long long flag[XK_KINDS][PAPR_XCHG_MAX_RANKS];
    // abort[q] != 0: rank q gave up waiting (timeout) during the exchange with that sequence number of
    // kind XK_PRE; every later wait on any rank fails too, so all ranks report the same outcome
    unsigned long long abort[PAPR_XCHG_MAX_RANKS];
    PaprXchgSlot slot[PAPR_XCHG_MAX_RANKS];
};

struct PaprPeers {
    PaprXchg *win[PAPR_XCHG_MAX_RANKS]; // win[rank] = the local window, the others are peer mappings
    int rank, world;
    unsigned long long timeout_ns;      // how long a kernel waits for a peer's publication
};

// launchers (papr_kernels.cu)
struct PaprTables {
    const double *pow10;     // [PAPR_MAX_LEVELS] pow(10, x_j) with the host libm, x_j per papr.c:139 / 170-172
    const double *ratio_min; // [PAPR_MAX_LEVELS] least peak/avg ratio for which the reference runs level j
    int nlevels_max;
};

void papr_launch_scan(bool stats, bool hist, int grid, const PaprScanArgs &a, cudaStream_t s);
void papr_launch_stats_finalize(const PaprCtaPartial *wp, int nctas, unsigned long long n, PaprDevStats *out,
                                cudaStream_t s);
void papr_launch_finalize_levels(const PaprCtaPartial *wp, int nctas, unsigned long long n, PaprTables t, int graph,
                                 PaprDevStats *local, PaprDevStats *merged, PaprDevLevels *lv,
                                 unsigned long long *status_word, const PaprChainList *chain /* or NULL */,
                                 int *chain_report /* {status, why} or NULL */, cudaStream_t s);
void papr_launch_levels(const PaprDevStats *parts, int nparts, PaprTables t, int graph,
                        PaprDevStats *merged, PaprDevLevels *lv, unsigned long long *status_word, cudaStream_t s,
                        const PaprChainList *chain = nullptr, int *chain_report = nullptr);
void papr_launch_presample(const float *iq, unsigned long long nsamples, int stride, int grid,
                           double *cta_pre /* [grid*3] */, cudaStream_t s);
void papr_launch_presample_reduce(const double *cta_pre, int nctas, double *pre4, cudaStream_t s);
void papr_launch_plan_pred(const double *pre4, const double *cta_pre, int nctas, PaprTables t, float sigmas,
                           float bias, int fine_slots, PaprPlan *plan, unsigned *fine_base, cudaStream_t s,
                           int ncells_max = PAPR_NCELLS_MAX);
void papr_launch_plan_exact(const PaprDevLevels *lv, const PaprDevStats *merged, int fine_bytes_log2,
                            PaprPlan *plan, unsigned *fine_base, cudaStream_t s);
void papr_launch_zero_fine(const PaprPlan *plan, unsigned long long *g_fine, int grid, cudaStream_t s);
void papr_launch_resolve(const PaprPlan *plan, const unsigned *fine_base, const PaprDevLevels *lv,
                         const unsigned long long *g_hist,
                         const unsigned long long *g_fine, const unsigned long long *g_over,
                         unsigned long long *counts, unsigned long long *status_word, int grid, cudaStream_t s);
void papr_launch_bsearch(const float *iq, unsigned long long nsamples, const PaprPlan *plan,
                         const PaprDevLevels *lv, unsigned long long *bhist, int grid, cudaStream_t s);
void papr_launch_bsearch_counts(const PaprPlan *plan, const PaprDevLevels *lv,
                                const unsigned long long *bhist, unsigned long long *counts,
                                cudaStream_t s);
void papr_launch_siggen(float *iq, unsigned long long first, unsigned long long nsamples,
                        unsigned long long seed, int grid, cudaStream_t s);
void papr_launch_find_nan(const float *iq, unsigned long long nsamples, unsigned long long first_index,
                          unsigned long long *out_idx, int grid, cudaStream_t s);
void papr_launch_tilesum(const float *iq, unsigned long long nsamples, double *tile_sum, int grid, cudaStream_t s);
void papr_launch_seqsum(const float *iq, unsigned long long nsamples, const short *tile_code,
                        void *tile_run /* {double e0, e1}[ntiles] */, int grid, cudaStream_t s);
int papr_seqsum_configure(void);
void papr_launch_plan_pred_x(const double *cta_pre, int nctas, PaprTables t, float sigmas, float bias, int fine_slots,
                             PaprPlan *plan, unsigned *fine_base, PaprPeers pp, unsigned long long seq, cudaStream_t s,
                             int ncells_max = PAPR_NCELLS_MAX);
// sharded: the statistics exchange + levels (papr_finalize.cuh: finalize_levels_x_body)
struct PaprFinalizeXArgs {
    const PaprCtaPartial *wp;
    int nctas;
    unsigned long long n;
    PaprDevStats *local;
    PaprTables tb;
    int graph;
    PaprDevStats *merged;
    PaprDevLevels *lv;
    unsigned long long *status_word;
    PaprPlan *plan;
    PaprPeers pp;
    unsigned long long seq;
    double bias; // test hook (1.0): scales the sum the levels are derived from
};

void papr_launch_finalize_levels_x(const PaprFinalizeXArgs &a, cudaStream_t s);
void papr_launch_counts_x(unsigned long long *counts, const PaprDevLevels *lv, PaprPlan *plan, PaprPeers pp,
                          unsigned long long seq, cudaStream_t s);
void papr_launch_tr(float *x, int nsym, int n, const float *kernel, const int *tone, int ntones, float vclip, int iterations,
                    float amax, float *r_out, int *iters_out, int grid, cudaStream_t s);
int papr_scan_tma_configure(void);
void papr_launch_scan_tma(const void *tensor_map /* CUtensorMap */, int grid, const PaprScanArgs &a, const PaprExactArgs &x,
                          cudaStream_t s);
void papr_launch_xt_compose(const PaprTileRun *tile_run, const int *tile_code, unsigned ntiles,
                            const PaprTileRun *multi_tile, PaprSuperRec *super, PaprSuperRec *hyper, int grid, cudaStream_t s,
                            unsigned long long *zero_word = nullptr /* status word of the counts that follow */,
                            const PaprFinalizeXArgs *fx = nullptr /* sharded: one extra CTA runs the statistics exchange */);
// single shard: chain (CTA 0) side by side with finalize + levels + counts from the fixed-order sum (the other CTAs)
struct PaprEpilogueArgs {
    const PaprCtaPartial *wp;
    int nctas;
    unsigned long long n;
    PaprTables tb;
    int graph;
    double bias;                       // test hook (1.0): scales the sum the speculative levels are derived from
    PaprDevStats *local, *merged;
    PaprDevLevels *lv;
    const PaprPlan *plan;
    const unsigned *fine_base;
    const unsigned long long *g_hist, *g_fine, *g_over;
    unsigned long long *counts, *status_word;
    unsigned *done;                    // sharded: counting CTAs that have finished (the last one exchanges the counts)
    int *chain_report;                 // {XT_* status, why}
    double *chain_exact;               // the chained sequential sum (valid when status == XT_OK)
};
// sharded: the chain with its exchange (CTA 0) side by side with this shard's level counts against a.lv (the other CTAs)
void papr_launch_xt_epilogue_x(const PaprSuperRec *hyper, const PaprSuperRec *super, const PaprTileRun *tile_run,
                               const int *tile_code, const PaprTileRun *multi, const PaprTileRun *multi_tile, unsigned ntiles,
                               const float *iq, unsigned long long nsamples, PaprChainList *out, PaprPlan *plan, PaprPeers pp,
                               unsigned long long seq, const PaprEpilogueArgs &a, int grid, cudaStream_t s,
                               int decline /* this rank has no runs: publish an empty list marked XT_FALLBACK */,
                               unsigned long long seq_counts /* of the counts exchange in the kernel's tail */);
void papr_launch_xt_epilogue(const PaprSuperRec *hyper, const PaprSuperRec *super, const PaprTileRun *tile_run, const int *tile_code,
                             const PaprTileRun *multi, const PaprTileRun *multi_tile, unsigned ntiles, const float *iq,
                             unsigned long long nsamples, PaprChainList *out, const PaprEpilogueArgs &a, int grid, cudaStream_t s);
int papr_scan_smem_bytes(bool hist);
int papr_scan_configure(void); // sets the dynamic shared-memory attributes once per device
