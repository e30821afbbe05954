// papr_engine.cu — the C ABI (include/papr_b200.h): engine lifetime, the device-resident analysis,
// the host-side analysis (files, pinned and pageable buffers streamed in chunks with the statistics
// pass and the sequential-sum emulation running under the H2D shadow; re-streaming for captures larger
// than HBM), the per-stage entry points for sharded callers, the sharded analysis with its exchanges
// fused into the kernels over peer memory, and the single-process multi-GPU driver.
//
// One analysis is a single stream-ordered chain of launches with ONE host synchronisation at the
// end: the reference's scalar epilogue (papr.c:131-141 / 164-173) is evaluated on the device from
// host-libm tables (papr_levels_kernel), then re-evaluated on the host with the real libm and
// compared bit for bit before anything is reported.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <ctype.h>
#include <errno.h>
#include <pthread.h>
#include <sched.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "papr_device.cuh"
#include "papr_host.h"

typedef unsigned long long u64;

namespace {

constexpr int kMaxStages = 34;             // pinned staging slots for pageable / file sources (threads + 2)
constexpr int kRing = 4;                   // device chunk buffers of the re-streaming (capture > HBM) mode
constexpr int kSeqLagRing = 3;             // chunks between a chunk's tile sums and its tile runs (< kRing) ...
constexpr int kSeqLagResident = 6;         // ... and when nothing limits how far the copies run ahead (< kSeqEvents)
constexpr int kSeqEvents = 8;
constexpr size_t kTileBytes = (size_t)PAPR_SEQ_TILE * 8; // chunks are whole tiles of the exact-sum emulation
constexpr size_t kPieceBytes = 4u << 20;   // staging granularity: small enough to still sit in the host's last-level
                                           // cache when the DMA engine reads it back, and to keep the pinned
                                           // allocation (0.6 ms per MiB on the boxes measured) short
constexpr u64 kMaxLaunchSamples = 1ull << 31;  // per scan launch (32-bit sample offsets inside a launch)
constexpr int kLevels1dB = 256;            // table sizes: PAPR < 192.7 dB (papr_b200.h)
constexpr int kLevelsGraph = PAPR_MAX_LEVELS;

constexpr int kMaxGrid = 1024;

// device scratch that ONE memset clears before an analysis (all-zero = the reference's initial state)
struct DevWork {
    u64 hist[PAPR_NCELLS_MAX];
    u64 over;
    u64 bhist[PAPR_MAX_LEVELS + 1];
    PaprCtaPartial wp[kMaxGrid];
};

// device results that ONE D2H copy fetches; HostOut (pinned) mirrors it
struct DevOut {
    PaprDevStats local;   // this shard
    PaprDevStats merged;  // all shards (single-shard analysis: == local)
    PaprPlan plan;
    u64 counts[PAPR_MAX_LEVELS + 1]; // [PAPR_MAX_LEVELS] = status word (RES_* bits, summed over ranks)
    PaprDevLevels lv;
    int chain[2];                    // {XT_* status, why} of the sequential sum chained on the device
    double chain_exact;              // ... and the sum itself (single-shard epilogue; valid when status == XT_OK)
};

struct HostOut {
    DevOut o;
    u64 nan_idx;
    float nan_sample[2];
    double pre4[4];
};

std::string g_create_error;

// where the capture's bytes come from: a memory image (pinned or pageable) or a regular file, read
// with pread() straight into the pinned staging buffers (no page faults, no mapping to tear down)
struct HostSource {
    const unsigned char *img = nullptr;
    int fd = -1;
    int dfd = -1;  // the same file opened with O_DIRECT (tunable "o_direct"): aligned pieces bypass the page cache
    u64 base = 0;  // file offset of byte 0 of this source (shards of one file)
    u64 bytes = 0;
    bool pinned = false;
    HostSource slice(u64 off, u64 len) const
    {
        HostSource s = *this;
        if (s.img) s.img += off; else s.base += off;
        s.bytes = len;
        return s;
    }
    bool read(void *dst, u64 off, size_t len) const
    {
        if (img) { memcpy(dst, img + off, len); return true; }
        off += base;
        char *d = (char *)dst;
        if (dfd >= 0 && ((off | len | (u64)(uintptr_t)dst) & 4095) == 0) { // straight from the device into the pinned slot
            u64 o = off;
            size_t l = len;
            char *q = d;
            while (l) {
                ssize_t got = pread(dfd, q, l, (off_t)o);
                if (got < 0 && errno == EINTR) continue;
                if (got <= 0 || (got & 4095)) break; // not supported here after all / short read: the buffered descriptor finishes
                q += got; o += (u64)got; l -= (size_t)got;
            }
            if (l == 0) return true;
            d = q; off = o; len = l;
        }
        while (len) {
            ssize_t got = pread(fd, d, len, (off_t)off);
            if (got < 0 && errno == EINTR) continue;
            if (got <= 0) return false; // I/O error, or the file shrank under us
            d += got; off += (u64)got; len -= (size_t)got;
        }
        return true;
    }
};

// Fills the pinned staging slots of a pageable / file source ahead of the H2D copies: every helper
// thread takes the next whole chunk, waits until the slot's previous tenant has crossed PCIe, and
// reads the chunk into it (memcpy or pread); the main thread consumes the chunks in order.  No
// per-chunk fork/join: the helpers run free, the slots bound how far they get ahead.
class ChunkFeeder {
public:
    ChunkFeeder(int device, const HostSource &src, u64 nchunks, u64 chunk_bytes, u64 total_bytes, int threads,
                void *const *slots, const cudaEvent_t *slot_done, int nslots)
        : device_(device), src_(src), nchunks_(nchunks), chunk_bytes_(chunk_bytes), total_(total_bytes), slots_(slots),
          slot_done_(slot_done), nslots_(nslots), state_(nchunks, 0)
    {
        const int n = (int)std::min<u64>((u64)std::max(1, threads), std::max<u64>(nchunks, 1));
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~ChunkFeeder()
    {
        {
            std::lock_guard<std::mutex> l(m_);
            abort_ = true;
        }
        cv_slot_.notify_all();
        for (auto &t : workers_) t.join();
    }
    // blocks until chunk c sits in its slot; nullptr if reading it failed
    const void *wait_filled(u64 c)
    {
        std::unique_lock<std::mutex> l(m_);
        cv_fill_.wait(l, [&] { return state_[c] != 0; });
        return state_[c] == 1 ? slots_[c % nslots_] : nullptr;
    }
    // the H2D copy of chunk c has been enqueued and slot_done[c % nslots] recorded behind it
    void issued(u64 c)
    {
        {
            std::lock_guard<std::mutex> l(m_);
            issued_ = c + 1;
        }
        cv_slot_.notify_all();
    }

private:
    void run()
    {
        cudaSetDevice(device_);
        for (;;) {
            u64 c;
            {
                std::unique_lock<std::mutex> l(m_);
                if (abort_ || next_ >= nchunks_) return;
                c = next_++;
                if (c >= (u64)nslots_) { // slot still holds chunk c - nslots until that one has been issued ...
                    cv_slot_.wait(l, [&] { return abort_ || issued_ > c - nslots_; });
                    if (abort_) return;
                }
            }
            // ... and that copy (or, for the first chunks, whatever an earlier pass left in the slot) has
            // crossed PCIe; a never-recorded event returns at once
            cudaEventSynchronize(slot_done_[c % nslots_]);
            const u64 off = c * chunk_bytes_, len = std::min(chunk_bytes_, total_ - off);
            const bool ok = src_.read(slots_[c % nslots_], off, len);
            {
                std::lock_guard<std::mutex> l(m_);
                state_[c] = ok ? 1 : 2;
            }
            cv_fill_.notify_all();
        }
    }
    int device_;
    const HostSource &src_;
    u64 nchunks_, chunk_bytes_, total_;
    void *const *slots_;
    const cudaEvent_t *slot_done_;
    int nslots_;
    std::vector<char> state_; // 0 = not yet, 1 = filled, 2 = read error
    u64 next_ = 0, issued_ = 0;
    bool abort_ = false;
    std::mutex m_;
    std::condition_variable cv_fill_, cv_slot_;
    std::vector<std::thread> workers_;
};

} // namespace

struct papr_engine {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    std::string err;
    // tunables
    int mode = PAPR_MODE_AUTO;
    int presample_stride = 128; // upper bound; see presample_stride_for()
    float window_sigmas = 5.0f;
    float predict_bias = 1.0f; // test hook: scales the predicted mean of the fused mode (a wrong one must be caught)
    double epilogue_bias = 1.0; // test hook: scales the fixed-order sum the epilogue's speculative levels come from
    size_t chunk_bytes = 64u << 20; // H2D + kernel granularity; pageable / file sources are staged in 4 MiB pieces
    int staging_threads = -1;       // -1: hardware threads - 2, within [2, 16] (shared among the engines of a papr_multi)
    int host_share = 1;             // engines of this process that stage from the host at the same time
    int grid_per_sm = 1;
    u64 fused_min_samples = 1ull << 24;
    int fine_bytes_log2 = 26; // 64 MiB fine table
    int p2p_chain = -1;       // sharded in-kernel path: -1 = chain the sequential sum on the device if this shard qualifies
                              // (a rank whose shard does not declines inside the exchange: every rank then reports the
                              // fall-back); 0 = this rank always declines; 1 = fail if this shard does not qualify
    int o_direct = 0;         // 1: regular files are also opened with O_DIRECT (NVMe -> pinned staging, no page cache)
    int exact_sum = -1;       // != 0 (default): the reference's sequential double sum, bit for bit, on every path; 0: off
    // exact sequential-sum scratch (grown on demand)
    u64 seq_tiles = 0;
    double *d_tile_sum = nullptr, *h_tile_sum = nullptr;
    short *d_tile_code = nullptr, *h_tile_code = nullptr;
    double *d_tile_run = nullptr, *h_tile_run = nullptr; // {increment for an even, an odd entry state} per tile
    float *h_tile_data = nullptr;
    unsigned seq_dirty = 0;
    const float *seq_sums_ptr = nullptr; // h_tile_sum currently holds the tile sums of this resident shard ...
    u64 seq_sums_n = 0;                  // ... of this many samples (computed under the H2D shadow)
    // sequential sum inside the fused scan (papr_exact.cu): tile runs, super-tile records, multi-tile log, chain
    u64 xt_tiles = 0;
    unsigned xt_multi_cap = 8192;
    PaprTileRun *d_xt_run = nullptr, *d_xt_multi = nullptr, *d_xt_multi_tile = nullptr;
    int *d_xt_code = nullptr;
    PaprSuperRec *d_xt_super = nullptr, *d_xt_hyper = nullptr;
    unsigned *d_xt_multi_count = nullptr;
    PaprChainList *d_xt_chain = nullptr;
    PaprDevStats *d_xt_parts = nullptr;  // sharded: every rank's pass-1 state, in rank order
    unsigned *d_epi_done = nullptr;      // sharded epilogue: counting CTAs finished (self-resetting)
    int xt_status = -1, xt_why = 0;      // of the last analysis (-1: the device chain did not run)
    // device work buffers
    int grid = 0;
    DevWork *d_work = nullptr;
    DevOut *d_out = nullptr;
    u64 *d_nan_idx = nullptr;
    unsigned *d_fine_base = nullptr;
    u64 *d_fine = nullptr;
    double *d_pre_cta = nullptr, *d_pre4 = nullptr;
    double *d_tables = nullptr; // pow10[2][MAX] | ratio_min[2][MAX]
    HostOut *h_out = nullptr;
    // host path
    float *d_buf = nullptr;
    size_t d_buf_bytes = 0;
    void *h_stage[kMaxStages] = {};
    size_t h_stage_bytes = 0;
    int h_stage_slots = 0;
    cudaEvent_t stage_done[kMaxStages] = {};
    cudaEvent_t chunk_ready = nullptr;
    // re-streaming mode (capture larger than the resident budget): ring of device chunk buffers
    float *d_ring = nullptr;
    size_t d_ring_chunk = 0;
    cudaEvent_t ring_free[kRing] = {};
    cudaEvent_t seq_ev[kSeqEvents] = {};
    u64 max_resident_bytes = 0; // 0 = whatever cudaMemGetInfo leaves after a 1 GiB reserve
    bool streamed = false;      // the last host-path analysis ran in re-streaming mode
    // peer-memory exchange (one process per GPU): own window + cudaIpc mappings of the other ranks'
    PaprXchg *d_xchg = nullptr;
    PaprPeers peers = {};
    bool xchg_attached = false;
    u64 xseq[XK_KINDS] = {0, 0, 0, 0};
    double xchg_timeout_s = 30.0; // how long a kernel waits for a peer's publication before every rank gives up
    // timing / accounting
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_scan[4] = {nullptr, nullptr, nullptr, nullptr};
    int scan_pairs = 0;
    unsigned launches = 0;
    u64 h2d = 0, d2h = 0;
    int shard_mode = PAPR_MODE_TWO_PASS; // of the stream-ordered shard stages in flight

    PaprTables tables(int graph) const
    {
        PaprTables t;
        t.pow10 = d_tables + (graph ? PAPR_MAX_LEVELS : 0);
        t.ratio_min = d_tables + 2 * PAPR_MAX_LEVELS + (graph ? PAPR_MAX_LEVELS : 0);
        t.nlevels_max = graph ? kLevelsGraph : kLevels1dB;
        return t;
    }
};

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            e->err = std::string(#call) + ": " + cudaGetErrorString(e_);                      \
            return PAPR_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

static int fail(papr_engine *e, int code, const std::string &msg)
{
    e->err = msg;
    return code;
}

// everything except the two big arrays (which are valid up to nlevels only): the statistics head AND the
// diagnostics tail, so that a caller's stack struct never leaks garbage into fused_miss / the counters
static void reset_result(papr_result *out)
{
    memset(out, 0, offsetof(papr_result, level));
    memset(&out->mode_used, 0, sizeof(papr_result) - offsetof(papr_result, mode_used));
}

// ------------------------------------------------------------------------------------------------
// lifetime
// ------------------------------------------------------------------------------------------------
extern "C" int papr_abi_version(void) { return PAPR_B200_ABI_VERSION; }

extern "C" const char *papr_last_error(const papr_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" void *papr_engine_stream(papr_engine *e) { return e ? (void *)e->stream : nullptr; }

static int engine_init(papr_engine *e, int device)
{
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (count == 0) return fail(e, PAPR_ERR_CUDA, "no CUDA device");
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= count) return fail(e, PAPR_ERR_ARG, "device index out of range");
    CU(cudaSetDevice(device));
    e->device = device;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(e, PAPR_ERR_CUDA, "this library is built for sm_100a (B200) only");
    e->num_sms = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    if (papr_scan_configure() != 0 || papr_seqsum_configure() != 0 || papr_scan_tma_configure() != 0)
        return fail(e, PAPR_ERR_CUDA, "cudaFuncSetAttribute(shared memory) failed");
    e->grid = e->num_sms * e->grid_per_sm;
    if (e->grid > kMaxGrid) return fail(e, PAPR_ERR_CUDA, "more SMs than this build supports");
    CU(cudaMalloc(&e->d_work, sizeof(DevWork)));
    CU(cudaMalloc(&e->d_out, sizeof(DevOut)));
    CU(cudaMemset(e->d_out, 0, sizeof(DevOut)));
    CU(cudaMalloc(&e->d_nan_idx, sizeof(u64)));
    CU(cudaMalloc(&e->d_fine_base, sizeof(unsigned) * PAPR_NCELLS_MAX));
    CU(cudaMalloc(&e->d_fine, (size_t)1 << e->fine_bytes_log2));
    CU(cudaMalloc(&e->d_pre_cta, sizeof(double) * 3 * e->grid));
    CU(cudaMalloc(&e->d_pre4, sizeof(double) * 4));
    CU(cudaMalloc(&e->d_tables, sizeof(double) * 4 * PAPR_MAX_LEVELS));
    CU(cudaHostAlloc(&e->h_out, sizeof(HostOut), cudaHostAllocDefault));
    CU(cudaMalloc(&e->d_xt_multi, sizeof(PaprTileRun) * (size_t)e->xt_multi_cap * XT_MAX_CAND * XT_TILE_BATCHES));
    CU(cudaMalloc(&e->d_xt_multi_tile, sizeof(PaprTileRun) * (size_t)e->xt_multi_cap * XT_MAX_CAND));
    CU(cudaMalloc(&e->d_xt_multi_count, sizeof(unsigned)));
    CU(cudaMalloc(&e->d_xt_chain, sizeof(PaprChainList)));
    CU(cudaMemset(e->d_xt_chain, 0, sizeof(PaprChainList)));
    CU(cudaMalloc(&e->d_xt_parts, sizeof(PaprDevStats) * PAPR_XCHG_MAX_RANKS));
    CU(cudaMalloc(&e->d_epi_done, sizeof(unsigned)));
    CU(cudaMemset(e->d_epi_done, 0, sizeof(unsigned)));
    {
        std::vector<double> t(4 * PAPR_MAX_LEVELS, INFINITY);
        papr_host_build_tables(0, kLevels1dB, &t[0], &t[2 * PAPR_MAX_LEVELS]);
        papr_host_build_tables(1, kLevelsGraph, &t[PAPR_MAX_LEVELS], &t[3 * PAPR_MAX_LEVELS]);
        CU(cudaMemcpy(e->d_tables, t.data(), sizeof(double) * t.size(), cudaMemcpyHostToDevice));
    }
    CU(cudaEventCreate(&e->ev_begin));
    CU(cudaEventCreate(&e->ev_end));
    for (auto &ev : e->ev_scan) CU(cudaEventCreate(&ev));
    for (auto &ev : e->stage_done) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : e->ring_free) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (auto &ev : e->seq_ev) CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e->chunk_ready, cudaEventDisableTiming));
    return PAPR_OK;
}

extern "C" int papr_engine_create(int device, papr_engine **out)
{
    if (!out) return PAPR_ERR_ARG;
    *out = nullptr;
    papr_engine *e = new papr_engine();
    int rc = engine_init(e, device);
    if (rc != PAPR_OK) {
        g_create_error = e->err;
        papr_engine_destroy(e);
        return rc;
    }
    *out = e;
    return PAPR_OK;
}

extern "C" int papr_xchg_detach(papr_engine *e);

extern "C" void papr_engine_destroy(papr_engine *e)
{
    if (!e) return;
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_work); cudaFree(e->d_out); cudaFree(e->d_nan_idx); cudaFree(e->d_fine_base); cudaFree(e->d_fine);
    cudaFree(e->d_pre_cta); cudaFree(e->d_tile_sum); cudaFree(e->d_tile_code); cudaFree(e->d_tile_run);
    if (e->h_tile_sum) cudaFreeHost(e->h_tile_sum);
    if (e->h_tile_code) cudaFreeHost(e->h_tile_code);
    if (e->h_tile_run) cudaFreeHost(e->h_tile_run);
    if (e->h_tile_data) cudaFreeHost(e->h_tile_data);
    cudaFree(e->d_pre4); cudaFree(e->d_tables); cudaFree(e->d_buf);
    cudaFree(e->d_xt_run); cudaFree(e->d_xt_code); cudaFree(e->d_xt_super); cudaFree(e->d_xt_hyper); cudaFree(e->d_xt_multi); cudaFree(e->d_xt_multi_tile);
    cudaFree(e->d_xt_multi_count); cudaFree(e->d_xt_chain); cudaFree(e->d_xt_parts); cudaFree(e->d_epi_done);
    if (e->h_out) cudaFreeHost(e->h_out);
    if (e->h_stage[0]) cudaFreeHost(e->h_stage[0]);
    for (auto &ev : e->stage_done) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->ring_free) if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->seq_ev) if (ev) cudaEventDestroy(ev);
    cudaFree(e->d_ring);
    papr_xchg_detach(e);
    cudaFree(e->d_xchg);
    for (auto &ev : e->ev_scan) if (ev) cudaEventDestroy(ev);
    if (e->chunk_ready) cudaEventDestroy(e->chunk_ready);
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    delete e;
}

extern "C" int papr_engine_set(papr_engine *e, const char *name, double v)
{
    if (!e || !name) return PAPR_ERR_ARG;
    std::string n(name);
    if (n == "mode") e->mode = (int)v;
    else if (n == "presample_stride") e->presample_stride = std::max(1, (int)v);
    else if (n == "window_sigmas") e->window_sigmas = (float)v;
    else if (n == "predict_bias") e->predict_bias = (float)v;
    else if (n == "epilogue_bias") e->epilogue_bias = v;
    else if (n == "chunk_bytes") e->chunk_bytes = std::max<size_t>(1u << 20, ((size_t)v) & ~(kTileBytes - 1));
    else if (n == "staging_threads") e->staging_threads = std::max(-1, (int)v);
    else if (n == "max_resident_bytes") e->max_resident_bytes = (u64)v;
    else if (n == "fused_min_samples") e->fused_min_samples = (u64)v;
    else if (n == "exact_sum") e->exact_sum = (int)v;
    else if (n == "o_direct") e->o_direct = v != 0;
    else if (n == "p2p_chain") e->p2p_chain = (int)v;
    else if (n == "fine_bytes_log2") // <= the allocation, >= two slots of the finest cell size (u64 per float32 value)
        e->fine_bytes_log2 = std::min(26, std::max(3 + PAPR_SH_MIN + 1, (int)v));
    else if (n == "xchg_timeout_s") {
        e->xchg_timeout_s = std::min(3600.0, std::max(0.001, v));
        e->peers.timeout_ns = (u64)(e->xchg_timeout_s * 1e9);
    }
    else if (n == "grid_per_sm") return fail(e, PAPR_ERR_ARG, "grid_per_sm is fixed at engine creation");
    else return fail(e, PAPR_ERR_ARG, "unknown tunable: " + n);
    return PAPR_OK;
}

// ------------------------------------------------------------------------------------------------
// launch chains (everything below only enqueues on e->stream)
// ------------------------------------------------------------------------------------------------
static int enqueue_reset(papr_engine *e)
{
    CU(cudaMemsetAsync(e->d_work, 0, sizeof(DevWork), e->stream));
    return PAPR_OK;
}

static void scan_args(papr_engine *e, PaprScanArgs &a, const float *d_iq, u64 n, u64 first)
{
    a.iq = d_iq;
    a.nsamples = n;
    a.first_index = first;
    a.wp = e->d_work->wp;
    a.plan = &e->d_out->plan;
    a.fine_base = e->d_fine_base;
    a.g_hist = e->d_work->hist;
    a.g_fine = e->d_fine;
    a.g_over = &e->d_work->over;
}

// one pass over [d_iq, d_iq + 2n) split into launches of < 2^32 samples at batch-aligned cuts
static int enqueue_scan(papr_engine *e, bool stats, bool hist, const float *d_iq, u64 n, u64 first, bool timed)
{
    if (timed && e->scan_pairs < 2) CU(cudaEventRecord(e->ev_scan[2 * e->scan_pairs], e->stream));
    for (u64 off = 0; off < n || (off == 0 && n == 0 && stats); off += kMaxLaunchSamples) {
        u64 m = std::min(kMaxLaunchSamples, n - off);
        if (m == 0) break;
        PaprScanArgs a;
        scan_args(e, a, d_iq + 2 * off, m, first + off);
        papr_launch_scan(stats, hist, e->grid, a, e->stream);
        e->launches += 1;
    }
    if (timed && e->scan_pairs < 2) {
        CU(cudaEventRecord(e->ev_scan[2 * e->scan_pairs + 1], e->stream));
        e->scan_pairs++;
    }
    CU(cudaGetLastError());
    return PAPR_OK;
}

// ---- the TMA-fed fused scan with the sequential sum inside (papr_exact.cu) ------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static TensorMapEncodeFn tensor_map_encoder()
{
    static TensorMapEncodeFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
        return (TensorMapEncodeFn)p;
    }();
    return fn;
}

static int ensure_xt_buffers(papr_engine *e, u64 ntiles)
{
    if (ntiles <= e->xt_tiles) return PAPR_OK;
    cudaFree(e->d_xt_run); cudaFree(e->d_xt_code); cudaFree(e->d_xt_super); cudaFree(e->d_xt_hyper);
    e->d_xt_run = nullptr; e->d_xt_code = nullptr; e->d_xt_super = nullptr; e->d_xt_hyper = nullptr;
    e->xt_tiles = 0;
    const u64 cap = std::max<u64>(ntiles, 1u << 16);
    CU(cudaMalloc(&e->d_xt_run, cap * sizeof(PaprTileRun)));
    CU(cudaMalloc(&e->d_xt_code, cap * sizeof(int)));
    CU(cudaMalloc(&e->d_xt_super, (cap / XT_SUPER_TILES + 1) * sizeof(PaprSuperRec)));
    CU(cudaMalloc(&e->d_xt_hyper, (cap / XT_SUPER_TILES / XT_HYPER_SUPERS + 1) * sizeof(PaprSuperRec)));
    e->xt_tiles = cap;
    return PAPR_OK;
}

// is the fused scan with the in-sweep sequential sum applicable to this shard?
static bool xt_applicable(const papr_engine *e, u64 n)
{
    return e->exact_sum != 0 && n >= 2 * XT_TILE_SAMPLES && tensor_map_encoder() != nullptr;
}

// the whole shard in launches of <= 2^31 samples (tile-aligned cuts); g_first = whole-capture index of sample 0
static int enqueue_scan_tma(papr_engine *e, const float *d_iq, u64 n, u64 first, bool timed)
{
    int rc;
    const u64 ntiles = (n + XT_TILE_SAMPLES - 1) / XT_TILE_SAMPLES;
    if ((rc = ensure_xt_buffers(e, ntiles))) return rc;
    CU(cudaMemsetAsync(e->d_xt_multi_count, 0, sizeof(unsigned), e->stream));
    if (timed && e->scan_pairs < 2) CU(cudaEventRecord(e->ev_scan[2 * e->scan_pairs], e->stream));
    for (u64 off = 0, m = 0; off < n; off += m) {
        m = std::min(kMaxLaunchSamples, n - off);
        if (n - off - m < 2 * XT_TILE_SAMPLES) m = n - off; // never leave a sliver (a launch needs whole 128-byte rows)
        PaprScanArgs a;
        scan_args(e, a, d_iq + 2 * off, m, first + off);
        PaprExactArgs x;
        x.g_first = first + off;
        x.tile_base = (unsigned)(off / XT_TILE_SAMPLES);
        x.literal_tile0 = first + off == 0;
        x.tile_run = e->d_xt_run; x.tile_code = e->d_xt_code;
        x.multi = e->d_xt_multi; x.multi_tile = e->d_xt_multi_tile; x.multi_count = e->d_xt_multi_count; x.multi_cap = e->xt_multi_cap;
        CUtensorMap tm; // rows of 128 bytes (16 samples); the < 16 samples past the last full row are patched in by the kernel
        const cuuint64_t dims[2] = {32, std::max<u64>(m / 16, 1)};
        const cuuint64_t strides[1] = {128};
        const cuuint32_t box[2] = {32, 32}, es[2] = {1, 1}; // 32 rows x 128 B = one 4 KiB warp batch
        CUresult r = tensor_map_encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)(d_iq + 2 * off), dims, strides, box, es,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                          CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(e, PAPR_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
        papr_launch_scan_tma(&tm, e->grid, a, x, e->stream);
        e->launches += 1;
    }
    if (timed && e->scan_pairs < 2) {
        CU(cudaEventRecord(e->ev_scan[2 * e->scan_pairs + 1], e->stream));
        e->scan_pairs++;
    }
    CU(cudaGetLastError());
    return PAPR_OK;
}

static PaprEpilogueArgs epilogue_args(papr_engine *e, u64 n, int graph)
{
    PaprEpilogueArgs a;
    a.wp = e->d_work->wp; a.nctas = e->grid; a.n = n;
    a.tb = e->tables(graph); a.graph = graph; a.bias = e->epilogue_bias;
    a.local = &e->d_out->local; a.merged = &e->d_out->merged; a.lv = &e->d_out->lv;
    a.plan = &e->d_out->plan; a.fine_base = e->d_fine_base;
    a.g_hist = e->d_work->hist; a.g_fine = e->d_fine; a.g_over = &e->d_work->over;
    a.counts = e->d_out->counts; a.status_word = &e->d_out->counts[PAPR_MAX_LEVELS];
    a.chain_report = e->d_out->chain; a.chain_exact = &e->d_out->chain_exact;
    a.done = e->d_epi_done;
    return a;
}

// single shard, after the TMA-fed sweep: tile runs -> super-tile records, then ONE launch in which CTA 0 chains
// the sequential sum while the other CTAs do finalize + levels + counts from the fixed-order sum (papr_exact.cu:
// papr_xt_epilogue_kernel); the host checks the speculative levels against the ones the exact sum gives
static int enqueue_xt_epilogue(papr_engine *e, const float *d_iq, u64 n, int graph)
{
    const unsigned ntiles = (unsigned)((n + XT_TILE_SAMPLES - 1) / XT_TILE_SAMPLES);
    const unsigned nsuper = (ntiles + XT_SUPER_TILES - 1) / XT_SUPER_TILES;
    const unsigned nhyper = (nsuper + XT_HYPER_SUPERS - 1) / XT_HYPER_SUPERS;
    papr_launch_xt_compose(e->d_xt_run, e->d_xt_code, ntiles, e->d_xt_multi_tile, e->d_xt_super, e->d_xt_hyper,
                           (int)std::min<unsigned>(nhyper, (unsigned)e->num_sms * 2), e->stream,
                           &e->d_out->counts[PAPR_MAX_LEVELS]);
    const PaprEpilogueArgs a = epilogue_args(e, n, graph);
    papr_launch_xt_epilogue(e->d_xt_hyper, e->d_xt_super, e->d_xt_run, e->d_xt_code, e->d_xt_multi, e->d_xt_multi_tile, ntiles,
                            d_iq, n, e->d_xt_chain, a, e->num_sms, e->stream);
    e->launches += 2;
    CU(cudaGetLastError());
    return PAPR_OK;
}

// CTA partials -> local stats -> (single shard) merged stats, avg, L, levels on the device
static int enqueue_finalize_levels(papr_engine *e, u64 n, int graph, bool chained = false)
{
    papr_launch_finalize_levels(e->d_work->wp, e->grid, n, e->tables(graph), graph, &e->d_out->local, &e->d_out->merged,
                                &e->d_out->lv, &e->d_out->counts[PAPR_MAX_LEVELS], chained ? e->d_xt_chain : nullptr,
                                chained ? e->d_out->chain : nullptr, e->stream);
    e->launches += 1;
    CU(cudaGetLastError());
    return PAPR_OK;
}

static int enqueue_resolve(papr_engine *e)
{
    papr_launch_resolve(&e->d_out->plan, e->d_fine_base, &e->d_out->lv, e->d_work->hist, e->d_fine, &e->d_work->over,
                        e->d_out->counts, &e->d_out->counts[PAPR_MAX_LEVELS], e->num_sms * 2, e->stream);
    e->launches += 1;
    CU(cudaGetLastError());
    return PAPR_OK;
}

// exact-threshold CCDF pass (d_out->lv / d_out->merged already hold levels and peak), in three steps so
// that the middle one can run per chunk when the capture is re-streamed instead of resident
static int hist_begin(papr_engine *e)
{
    papr_launch_plan_exact(&e->d_out->lv, &e->d_out->merged, e->fine_bytes_log2, &e->d_out->plan, e->d_fine_base,
                           e->stream);
    papr_launch_zero_fine(&e->d_out->plan, e->d_fine, e->num_sms * 4, e->stream);
    e->launches += 2;
    CU(cudaGetLastError());
    return PAPR_OK;
}

static int hist_chunk(papr_engine *e, const float *d_iq, u64 n, bool timed)
{
    int rc = enqueue_scan(e, false, true, d_iq, n, 0, timed);
    if (rc) return rc;
    papr_launch_bsearch(d_iq, n, &e->d_out->plan, &e->d_out->lv, e->d_work->bhist, e->grid, e->stream); // no-op unless PLAN_BSEARCH
    e->launches += 1;
    CU(cudaGetLastError());
    return PAPR_OK;
}

static int hist_end(papr_engine *e)
{
    int rc = enqueue_resolve(e);
    if (rc) return rc;
    papr_launch_bsearch_counts(&e->d_out->plan, &e->d_out->lv, e->d_work->bhist, e->d_out->counts, e->stream);
    e->launches += 1;
    CU(cudaGetLastError());
    return PAPR_OK;
}

static int enqueue_hist_exact(papr_engine *e, const float *d_iq, u64 n, bool timed)
{
    int rc;
    if ((rc = hist_begin(e))) return rc;
    if ((rc = hist_chunk(e, d_iq, n, timed))) return rc;
    return hist_end(e);
}

static int enqueue_fetch(papr_engine *e)
{
    CU(cudaMemcpyAsync(&e->h_out->o, e->d_out, sizeof(DevOut), cudaMemcpyDeviceToHost, e->stream));
    e->d2h += sizeof(DevOut);
    return PAPR_OK;
}

static float bits_to_float(int b)
{
    float f;
    memcpy(&f, &b, 4);
    return f;
}

static void stats_to_host(const PaprDevStats &d, papr_stats *s)
{
    memset(s, 0, sizeof(*s));
    s->n = d.n;
    s->sum = d.sum;
    s->peak = bits_to_float(d.val[TR_PEAK]);
    s->peak_idx = d.idx[TR_PEAK];
    s->re_pos = bits_to_float(d.val[TR_RE_POS]);
    s->re_pos_idx = d.idx[TR_RE_POS];
    s->im_pos = bits_to_float(d.val[TR_IM_POS]);
    s->im_pos_idx = d.idx[TR_IM_POS];
    // the negative trackers hold magnitudes; 0 means "never below the 0.0 initial value" (papr.c:44-45)
    s->re_neg = d.val[TR_RE_NEG] ? -bits_to_float(d.val[TR_RE_NEG]) : 0.0f;
    s->re_neg_idx = d.idx[TR_RE_NEG];
    s->im_neg = d.val[TR_IM_NEG] ? -bits_to_float(d.val[TR_IM_NEG]) : 0.0f;
    s->im_neg_idx = d.idx[TR_IM_NEG];
    s->flags = d.flags;
}

static void stats_to_device(const papr_stats &s, PaprDevStats *d)
{
    auto fb = [](float f) { int b; memcpy(&b, &f, 4); return b; };
    d->sum = s.sum;
    d->n = s.n;
    d->val[TR_PEAK] = fb(s.peak);       d->idx[TR_PEAK] = s.peak_idx;
    d->val[TR_RE_POS] = fb(s.re_pos);   d->idx[TR_RE_POS] = s.re_pos_idx;
    d->val[TR_RE_NEG] = s.re_neg < 0 ? fb(-s.re_neg) : 0; d->idx[TR_RE_NEG] = s.re_neg_idx;
    d->val[TR_IM_POS] = fb(s.im_pos);   d->idx[TR_IM_POS] = s.im_pos_idx;
    d->val[TR_IM_NEG] = s.im_neg < 0 ? fb(-s.im_neg) : 0; d->idx[TR_IM_NEG] = s.im_neg_idx;
    d->flags = s.flags;
}

// ------------------------------------------------------------------------------------------------
// completion: one synchronisation, host re-evaluation of the epilogue, verification
// ------------------------------------------------------------------------------------------------
// Subsample density of the fused mode: the windows around the predicted thresholds shrink with
// 1/sqrt(subsample), and all of them must fit the fine table, so -g (0.1 dB spacing, ~10x the levels)
// wants a 4x larger subsample than the 1 dB mode.  Capped by the "presample_stride" tunable.
static int presample_stride_for(const papr_engine *e, u64 n, int graph)
{
    const u64 nbatch = n / PAPR_BATCH_SAMPLES, target = graph ? (1u << 18) : (1u << 16);
    return (int)std::max<u64>(1, std::min<u64>((u64)e->presample_stride, nbatch / target));
}

static void begin_analysis(papr_engine *e)
{
    e->scan_pairs = 0;
    e->launches = 0;
    e->h2d = e->d2h = 0;
    e->seq_sums_ptr = nullptr;
    cudaSetDevice(e->device);
}

// The reference's printed "nan" carries the sign of the first NaN that entered `sum` (x86 addsd /
// addss keep the destination operand's NaN: papr.c:103-104 compile to Q*Q + I*I and sum += value),
// while the GPU produces a canonical NaN.  Find that sample and restore the sign.
static int fix_nan_sign(papr_engine *e, const float *d_iq, u64 n, u64 first, papr_stats *st)
{
    if (!std::isnan(st->sum) || n == 0) return PAPR_OK;
    CU(cudaMemsetAsync(e->d_nan_idx, 0xff, sizeof(u64), e->stream));
    papr_launch_find_nan(d_iq, n, first, e->d_nan_idx, e->num_sms * 8, e->stream);
    CU(cudaMemcpyAsync(&e->h_out->nan_idx, e->d_nan_idx, sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    u64 k = e->h_out->nan_idx;
    if (k == ~0ull) return PAPR_OK; // Inf - Inf cannot occur (powers are >= 0); nothing to do
    CU(cudaMemcpyAsync(e->h_out->nan_sample, d_iq + 2 * (k - first), 8, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    float re = e->h_out->nan_sample[0], im = e->h_out->nan_sample[1];
    bool neg = std::isnan(im) ? std::signbit(im) : std::signbit(re);
    st->sum = std::copysign(std::fabs(st->sum), neg ? -1.0 : 1.0);
    e->launches += 1;
    e->d2h += 16;
    return PAPR_OK;
}

// Fill `out` from the fetched device results; non-zero = device level table disagrees with the host libm
// (or the fused pass missed): the caller must count again / run the exact CCDF pass with out->level[].
static int collect(papr_engine *e, int graph, papr_result *out, bool counts_valid)
{
    const DevOut *h = &e->h_out->o;
    papr_result_finish(out, graph);
    int redo = 0; // bit 0: the device's levels are not the host's; bit 1: the counts are invalid / a window was missed
    if (out->nlevels > 0) {
        bool same = h->lv.L == out->nlevels &&
                    memcmp(h->lv.level, out->level, sizeof(float) * (size_t)out->nlevels) == 0;
        if (!same) redo |= 1;
        if (!counts_valid || (h->counts[PAPR_MAX_LEVELS] != 0)) redo |= 2;
        if (!redo)
            for (int j = 0; j < out->nlevels; ++j) out->level_count[j] = (int64_t)h->counts[j];
    }
    return redo;
}

static int upload_levels(papr_engine *e, const float *level, int L, float peak)
{
    static thread_local PaprDevLevels lv;
    static thread_local PaprDevStats ms;
    memset(&ms, 0, sizeof(ms));
    memcpy(&ms.val[TR_PEAK], &peak, 4);
    lv.avg = 0; lv.ratio = 0; lv.L = L; lv.graph = 0; lv.pad[0] = lv.pad[1] = 0;
    memcpy(lv.level, level, sizeof(float) * (size_t)L);
    // pageable -> the runtime stages these synchronously, so the thread_local sources may be reused
    CU(cudaMemcpyAsync(&e->d_out->lv, &lv, offsetof(PaprDevLevels, level) + sizeof(float) * (size_t)L,
                       cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(&e->d_out->merged, &ms, sizeof(ms), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemsetAsync(&e->d_out->counts[PAPR_MAX_LEVELS], 0, sizeof(u64), e->stream));
    e->h2d += sizeof(float) * (size_t)L + sizeof(ms);
    return PAPR_OK;
}

// exact CCDF pass with host-provided levels; counts land in out (used for redo and papr_ccdf_device)
static int run_exact_ccdf(papr_engine *e, const float *d_iq, u64 n, const float *level, int L, float peak,
                          int64_t *level_count, bool accumulate)
{
    int rc = upload_levels(e, level, L, peak);
    if (rc) return rc;
    if ((rc = enqueue_reset(e))) return rc;
    rc = enqueue_hist_exact(e, d_iq, n, true);
    if (rc) return rc;
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaStreamSynchronize(e->stream));
    const DevOut *h = &e->h_out->o;
    if (h->counts[PAPR_MAX_LEVELS] != 0) return fail(e, PAPR_ERR_INTERNAL, "exact CCDF pass reported a miss");
    for (int j = 0; j < L; ++j) {
        if (accumulate) level_count[j] += (int64_t)h->counts[j];
        else level_count[j] = (int64_t)h->counts[j];
    }
    return PAPR_OK;
}

// The fused sweep's cells and fine table are still on the device: count again against out->level[] (the
// levels of the exact sum).  0 = out->level_count filled; 1 = a level falls outside its window (the caller
// runs the exact pass); < 0 = error.
static int recount(papr_engine *e, papr_result *out)
{
    int rc = upload_levels(e, out->level, out->nlevels, out->stats.peak);
    if (rc) return rc;
    if ((rc = enqueue_resolve(e))) return rc;
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaStreamSynchronize(e->stream));
    const DevOut *h = &e->h_out->o;
    if (h->counts[PAPR_MAX_LEVELS] != 0) return 1;
    for (int j = 0; j < out->nlevels; ++j) out->level_count[j] = (int64_t)h->counts[j];
    return 0;
}

static void finish_timing(papr_engine *e, papr_result *out)
{
    float ms = 0.f;
    out->scan_ms = 0.f;
    for (int p = 0; p < e->scan_pairs; ++p)
        if (cudaEventElapsedTime(&ms, e->ev_scan[2 * p], e->ev_scan[2 * p + 1]) == cudaSuccess) out->scan_ms += ms;
    if (cudaEventElapsedTime(&ms, e->ev_begin, e->ev_end) == cudaSuccess) out->device_ms = ms;
    out->kernel_launches = e->launches;
    out->h2d_bytes = e->h2d;
    out->d2h_bytes = e->d2h;
}

// ------------------------------------------------------------------------------------------------
// exact sequential sum (papr.c:104): see papr_seqsum_kernel.  Returns 1 if not applicable (NaN/Inf).
// ------------------------------------------------------------------------------------------------
static int ensure_seq_buffers(papr_engine *e, u64 ntiles)
{
    if (ntiles <= e->seq_tiles) return PAPR_OK;
    cudaFree(e->d_tile_sum); cudaFree(e->d_tile_code); cudaFree(e->d_tile_run);
    if (e->h_tile_sum) cudaFreeHost(e->h_tile_sum);
    if (e->h_tile_code) cudaFreeHost(e->h_tile_code);
    if (e->h_tile_run) cudaFreeHost(e->h_tile_run);
    e->seq_tiles = 0;
    u64 cap = std::max<u64>(ntiles, 4096);
    CU(cudaMalloc(&e->d_tile_sum, cap * sizeof(double)));
    CU(cudaMalloc(&e->d_tile_code, cap * sizeof(short)));
    CU(cudaMalloc(&e->d_tile_run, cap * 2 * sizeof(double)));
    CU(cudaHostAlloc(&e->h_tile_sum, cap * sizeof(double), cudaHostAllocDefault));
    CU(cudaHostAlloc(&e->h_tile_code, cap * sizeof(short), cudaHostAllocDefault));
    CU(cudaHostAlloc(&e->h_tile_run, cap * 2 * sizeof(double), cudaHostAllocDefault));
    if (!e->h_tile_data) CU(cudaHostAlloc(&e->h_tile_data, (size_t)PAPR_SEQ_TILE * 8, cudaHostAllocDefault));
    e->seq_tiles = cap;
    return PAPR_OK;
}

// phase 1: sums of this shard's tiles -> e->h_tile_sum (synchronises)
static int seq_tile_sums(papr_engine *e, const float *d_iq, u64 n)
{
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    int rc;
    if (ntiles && e->seq_sums_ptr == d_iq && e->seq_sums_n == n) return PAPR_OK; // already computed while streaming in
    if ((rc = ensure_seq_buffers(e, std::max<u64>(ntiles, 1)))) return rc;
    if (ntiles == 0) return PAPR_OK;
    papr_launch_tilesum(d_iq, n, e->d_tile_sum, e->num_sms * 2, e->stream);
    CU(cudaMemcpyAsync(e->h_tile_sum, e->d_tile_sum, ntiles * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->launches += 1;
    return PAPR_OK;
}

// phase 2 (host): place every tile in a binade, given the running sum `pre` before the shard.
// Returns 1 if a tile sum is NaN/Inf (the reference's sum is too: nothing to emulate).
static int seq_classify(const double *tile_sum, u64 ntiles, double *pre_io, short *code)
{
    // biased exponent field of a non-negative double; [2^k, 2^(k+1)) <=> field == k + 1023
    auto expo = [](double x) { u64 b; memcpy(&b, &x, 8); return (int)(b >> 52); };
    double pre = *pre_io;
    for (u64 t = 0; t < ntiles; ++t) {
        const double ts = tile_sum[t];
        if (!std::isfinite(ts)) return 1;
        short c = PAPR_SEQ_DIRTY;
        if (ts == 0.0) {
            c = PAPR_SEQ_ZERO;
        } else {
            // the exact running sum lies within 1e-9 (relative) of these approximate prefixes: the tile is
            // clean if both ends are normal numbers of the same binade
            const double lo = pre * (1.0 - 1e-9), hi = (pre + ts) * (1.0 + 1e-9);
            const int klo = expo(lo), khi = expo(hi);
            if (lo > 0.0 && klo != 0 && klo == khi) c = (short)(klo - 1023);
        }
        code[t] = c;
        pre += ts;
    }
    *pre_io = pre;
    return 0;
}

// phase 3: (D0, D1) of every clean tile -> e->h_tile_run (synchronises); codes in e->h_tile_code
static int seq_tile_runs(papr_engine *e, const float *d_iq, u64 n)
{
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    if (ntiles == 0) return PAPR_OK;
    CU(cudaMemcpyAsync(e->d_tile_code, e->h_tile_code, ntiles * sizeof(short), cudaMemcpyHostToDevice, e->stream));
    papr_launch_seqsum(d_iq, n, e->d_tile_code, e->d_tile_run, (int)std::min<u64>((u64)e->num_sms * 8, ntiles), e->stream);
    CU(cudaMemcpyAsync(e->h_tile_run, e->d_tile_run, ntiles * 2 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->launches += 1;
    e->d2h += ntiles * 24;
    e->h2d += ntiles * 2;
    return PAPR_OK;
}

// phase 4 (host): chain this shard's tiles onto the running state *s_io in file order; anything not
// provably inside its binade is replayed literally (real double adds on the tile's samples, which
// `fetch` brings to e->h_tile_data: a D2H copy for resident shards, a host read when re-streaming)
typedef std::function<int(u64 first, u64 cnt, float *dst)> TileFetch;

static int seq_chain(papr_engine *e, u64 n, double *s_io, const TileFetch &fetch)
{
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    double s = *s_io;
    for (u64 t = 0; t < ntiles; ++t) {
        const short code = e->h_tile_code[t];
        if (code == PAPR_SEQ_ZERO) continue;
        if (code != PAPR_SEQ_DIRTY) {
            u64 bits;
            memcpy(&bits, &s, 8);
            const unsigned ef = (unsigned)(bits >> 52); // sign + exponent field; s in [2^k, 2^(k+1)) <=> ef == k + 1023
            if (ef != 0 && ef == (unsigned)((int)code + 1023)) {
                // s = m * ulp with m's parity in bit 0; the tile's increment for that parity is a whole number
                // of ulps, so the add is exact as long as the result stays inside the binade - and a result
                // that leaves it (or a run whose accumulators left it: increment >= 2^k) is never below top
                const double inc = e->h_tile_run[2 * t + (bits & 1)];
                const u64 top_bits = (u64)(ef + 1) << 52;
                double top;
                memcpy(&top, &top_bits, 8);
                const double s2 = s + inc;
                if (inc >= 0.0 && s2 < top) {
                    s = s2;
                    continue;
                }
            }
        }
        const u64 first = t * PAPR_SEQ_TILE, cnt = std::min<u64>(PAPR_SEQ_TILE, n - first);
        int rc = fetch(first, cnt, e->h_tile_data);
        if (rc) return rc;
        s = papr_host_seq_add(s, e->h_tile_data, cnt);
        e->seq_dirty++;
    }
    *s_io = s;
    return PAPR_OK;
}

static int seq_chain_resident(papr_engine *e, const float *d_iq, u64 n, double *s_io)
{
    cudaSetDevice(e->device);
    return seq_chain(e, n, s_io, [&](u64 first, u64 cnt, float *dst) -> int {
        CU(cudaMemcpyAsync(dst, d_iq + 2 * first, cnt * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        e->d2h += cnt * 8;
        return PAPR_OK;
    });
}

static int exact_sequential_sum(papr_engine *e, const float *d_iq, u64 n, double *sum_out)
{
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    int rc;
    double pre = 0.0, s = 0.0;
    e->seq_dirty = 0;
    if ((rc = seq_tile_sums(e, d_iq, n))) return rc;
    if (seq_classify(e->h_tile_sum, ntiles, &pre, e->h_tile_code)) return 1;
    if ((rc = seq_tile_runs(e, d_iq, n))) return rc;
    if ((rc = seq_chain_resident(e, d_iq, n, &s))) return rc;
    *sum_out = s;
    return PAPR_OK;
}

// replace the tree sum of the local (== whole capture) state by the exact sequential one
static int apply_exact_sum(papr_engine *e, const float *d_iq, u64 n)
{
    double s = 0.0;
    int rc = exact_sequential_sum(e, d_iq, n, &s);
    if (rc < 0) return rc;
    if (rc == 0) {
        e->h_out->pre4[0] = s;
        CU(cudaMemcpyAsync(&e->d_out->local.sum, &e->h_out->pre4[0], sizeof(double), cudaMemcpyHostToDevice, e->stream));
    }
    return PAPR_OK;
}

// The same four phases as stages, for a capture sharded over processes (one rank per GPU): every rank
// prepares and computes its runs in parallel; only the chaining walks the ranks in index order.
extern "C" int papr_seqsum_prepare(papr_engine *e, const float *d_iq, uint64_t n, double *approx)
{
    if (!e || !approx || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    cudaSetDevice(e->device);
    int rc = seq_tile_sums(e, d_iq, n);
    if (rc) return rc;
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    double s = 0.0;
    for (u64 t = 0; t < ntiles; ++t) s += e->h_tile_sum[t];
    *approx = s;
    return PAPR_OK;
}

extern "C" int papr_seqsum_runs(papr_engine *e, const float *d_iq, uint64_t n, double pre)
{
    if (!e || (n && !d_iq)) return PAPR_ERR_ARG;
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    if (ntiles > e->seq_tiles) return fail(e, PAPR_ERR_ARG, "papr_seqsum_prepare must run first");
    cudaSetDevice(e->device);
    if (!std::isfinite(pre) || seq_classify(e->h_tile_sum, ntiles, &pre, e->h_tile_code)) return 1;
    return seq_tile_runs(e, d_iq, n);
}

extern "C" int papr_seqsum_chain(papr_engine *e, const float *d_iq, uint64_t n, double *state)
{
    if (!e || !state || (n && !d_iq)) return PAPR_ERR_ARG;
    const u64 ntiles = (n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    if (ntiles > e->seq_tiles) return fail(e, PAPR_ERR_ARG, "papr_seqsum_runs must run first");
    e->seq_dirty = 0;
    return seq_chain_resident(e, d_iq, n, state);
}

// finalize + levels, optionally with the exact sequential sum patched in between (two extra passes
// over the resident shard and two host synchronisations - negligible next to PCIe on the host path)
static int enqueue_finalize_levels_exact(papr_engine *e, const float *d_iq, u64 n, int graph, bool exact)
{
    if (!exact) return enqueue_finalize_levels(e, n, graph);
    papr_launch_stats_finalize(e->d_work->wp, e->grid, n, &e->d_out->local, e->stream);
    e->launches += 1;
    int rc = apply_exact_sum(e, d_iq, n);
    if (rc) return rc;
    papr_launch_levels(&e->d_out->local, 1, e->tables(graph), graph, &e->d_out->merged, &e->d_out->lv,
                       &e->d_out->counts[PAPR_MAX_LEVELS], e->stream);
    e->launches += 1;
    CU(cudaGetLastError());
    return PAPR_OK;
}

// ------------------------------------------------------------------------------------------------
// device-resident analysis
// ------------------------------------------------------------------------------------------------
extern "C" int papr_analyze_device(papr_engine *e, const float *d_iq, uint64_t n, int graph, papr_result *out)
{
    if (!e || !out || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    graph = graph ? 1 : 0;
    begin_analysis(e);
    reset_result(out);
    int mode = e->mode;
    const int stride = presample_stride_for(e, n, graph);
    if (mode == PAPR_MODE_AUTO) // fused pays off once the subsample is a small fraction of the shard
        mode = (n >= e->fused_min_samples && stride >= 8) ? PAPR_MODE_FUSED : PAPR_MODE_TWO_PASS;
    out->mode_used = mode;
    const bool exact = e->exact_sum != 0; // papr.c:104 bit for bit (tunable "exact_sum" = 0: any deterministic order)
    int rc;
    CU(cudaEventRecord(e->ev_begin, e->stream));
    if ((rc = enqueue_reset(e))) return rc;
    const bool chained = mode == PAPR_MODE_FUSED && xt_applicable(e, n); // sequential sum inside the sweep
    e->xt_status = -1;
    if (mode == PAPR_MODE_FUSED) {
        papr_launch_presample(d_iq, std::min<u64>(n, kMaxLaunchSamples * 2 - PAPR_BATCH_SAMPLES), stride,
                              e->grid, e->d_pre_cta, e->stream);
        papr_launch_plan_pred(nullptr, e->d_pre_cta, e->grid, e->tables(graph), e->window_sigmas, e->predict_bias,
                              1 << (e->fine_bytes_log2 - 3 - PAPR_SH_MIN), &e->d_out->plan, e->d_fine_base, e->stream,
                              chained ? XT_NCELLS : PAPR_NCELLS_MAX);
        papr_launch_zero_fine(&e->d_out->plan, e->d_fine, e->num_sms * 4, e->stream);
        e->launches += 3;
        if (chained) {
            if ((rc = enqueue_scan_tma(e, d_iq, n, 0, true))) return rc;
            if ((rc = enqueue_xt_epilogue(e, d_iq, n, graph))) return rc; // chain || finalize + levels + counts
        } else {
            if ((rc = enqueue_scan(e, true, true, d_iq, n, 0, true))) return rc;
            if ((rc = enqueue_finalize_levels_exact(e, d_iq, n, graph, exact))) return rc;
            if ((rc = enqueue_resolve(e))) return rc;
        }
    } else {
        if ((rc = enqueue_scan(e, true, false, d_iq, n, 0, true))) return rc;
        if ((rc = enqueue_finalize_levels_exact(e, d_iq, n, graph, exact))) return rc;
        if ((rc = enqueue_hist_exact(e, d_iq, n, true))) return rc;
    }
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaEventRecord(e->ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    stats_to_host(e->h_out->o.merged, &out->stats);
    if (chained) {
        e->xt_status = e->h_out->o.chain[0];
        e->xt_why = e->h_out->o.chain[1];
        if (e->xt_status == XT_OK) out->stats.sum = e->h_out->o.chain_exact; // replaces the fixed-order sum of the CTA partials
        if (e->xt_status == XT_FALLBACK) { // the device chain could not vouch for its sum: the two-sweep emulation
            double s = 0.0;
            if ((rc = exact_sequential_sum(e, d_iq, n, &s)) < 0) return rc;
            if (rc == 0) out->stats.sum = s; // levels are re-derived from it below; counts redone only if they move
        }
    }
    if (chained) out->sum_path = e->xt_status == XT_OK ? 1u : e->xt_status == XT_FALLBACK ? (2u | ((unsigned)e->xt_why << 8)) : 0u;
    else out->sum_path = exact && std::isfinite(out->stats.sum) ? 2u : 0u;
    if ((rc = fix_nan_sign(e, d_iq, n, 0, &out->stats))) return rc;
    if (const int redo = collect(e, graph, out, true)) {
        // chained: the counts were taken against levels derived from the fixed-order sum.  If those differ from
        // the levels of the exact sum (rare), the cells and the fine table still hold everything needed.
        bool recounted = false;
        if (chained && (redo & 1)) {
            if ((rc = recount(e, out)) < 0) return rc;
            recounted = rc == 0;
        }
        if (recounted) {
            out->fused_miss = 2;
        } else {
            out->fused_miss = mode == PAPR_MODE_FUSED;
            rc = run_exact_ccdf(e, d_iq, n, out->level, out->nlevels, out->stats.peak, out->level_count, false);
            if (rc) return rc;
        }
        CU(cudaEventRecord(e->ev_end, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    finish_timing(e, out);
    return PAPR_OK;
}

// ------------------------------------------------------------------------------------------------
// stages for sharded callers
// ------------------------------------------------------------------------------------------------
extern "C" int papr_stats_device(papr_engine *e, const float *d_iq, uint64_t n, uint64_t first, papr_stats *out)
{
    if (!e || !out || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    begin_analysis(e);
    int rc;
    if ((rc = enqueue_reset(e))) return rc;
    if ((rc = enqueue_scan(e, true, false, d_iq, n, first, true))) return rc;
    papr_launch_stats_finalize(e->d_work->wp, e->grid, n, &e->d_out->local, e->stream);
    e->launches += 1;
    CU(cudaMemcpyAsync(&e->h_out->o.local, &e->d_out->local, sizeof(PaprDevStats), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    stats_to_host(e->h_out->o.local, out);
    return fix_nan_sign(e, d_iq, n, first, out);
}

extern "C" int papr_ccdf_device(papr_engine *e, const float *d_iq, uint64_t n, const float *level, int L,
                                int64_t *level_count)
{
    if (!e || (n && !d_iq) || L < 0 || L > PAPR_MAX_LEVELS || (L && (!level || !level_count))) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    if (L == 0 || n == 0) return PAPR_OK;
    for (int j = 0; j < L; ++j)
        if (!(level[j] >= 0.0f) || (j && level[j] < level[j - 1]))
            return fail(e, PAPR_ERR_ARG, "levels must be non-negative and non-decreasing");
    begin_analysis(e);
    return run_exact_ccdf(e, d_iq, n, level, L, 0.0f, level_count, true);
}

extern "C" int papr_fused_presample(papr_engine *e, const float *d_iq, uint64_t n, int graph, double pre[4])
{
    if (!e || !pre || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    begin_analysis(e);
    papr_launch_presample(d_iq, std::min<u64>(n, kMaxLaunchSamples * 2 - PAPR_BATCH_SAMPLES),
                          presample_stride_for(e, n, graph ? 1 : 0), e->grid, e->d_pre_cta, e->stream);
    papr_launch_presample_reduce(e->d_pre_cta, e->grid, e->d_pre4, e->stream);
    CU(cudaMemcpyAsync(e->h_out->pre4, e->d_pre4, sizeof(double) * 4, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    memcpy(pre, e->h_out->pre4, sizeof(double) * 4);
    return PAPR_OK;
}

extern "C" int papr_fused_scan(papr_engine *e, const float *d_iq, uint64_t n, uint64_t first, const double pre[4],
                               int graph, papr_stats *out)
{
    if (!e || !pre || !out || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    graph = graph ? 1 : 0;
    begin_analysis(e);
    int rc;
    CU(cudaEventRecord(e->ev_begin, e->stream));
    if ((rc = enqueue_reset(e))) return rc;
    CU(cudaMemcpyAsync(e->d_pre4, pre, sizeof(double) * 4, cudaMemcpyHostToDevice, e->stream));
    papr_launch_plan_pred(e->d_pre4, nullptr, 0, e->tables(graph), e->window_sigmas, e->predict_bias,
                          1 << (e->fine_bytes_log2 - 3 - PAPR_SH_MIN), &e->d_out->plan, e->d_fine_base, e->stream);
    papr_launch_zero_fine(&e->d_out->plan, e->d_fine, e->num_sms * 4, e->stream);
    e->launches += 2;
    if ((rc = enqueue_scan(e, true, true, d_iq, n, first, true))) return rc;
    papr_launch_stats_finalize(e->d_work->wp, e->grid, n, &e->d_out->local, e->stream);
    e->launches += 1;
    CU(cudaMemcpyAsync(&e->h_out->o.local, &e->d_out->local, sizeof(PaprDevStats), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    stats_to_host(e->h_out->o.local, out);
    return fix_nan_sign(e, d_iq, n, first, out);
}

extern "C" int papr_fused_counts(papr_engine *e, const papr_stats *merged, int graph, int64_t *level_count)
{
    if (!e || !merged || !level_count) return PAPR_ERR_ARG;
    graph = graph ? 1 : 0;
    static thread_local papr_result r;
    r.stats = *merged;
    papr_result_finish(&r, graph);
    if (r.nlevels == 0) return PAPR_OK;
    PaprDevStats ms;
    stats_to_device(*merged, &ms);
    CU(cudaMemcpyAsync(&e->d_out->local, &ms, sizeof(ms), cudaMemcpyHostToDevice, e->stream));
    papr_launch_levels(&e->d_out->local, 1, e->tables(graph), graph, &e->d_out->merged, &e->d_out->lv,
                       &e->d_out->counts[PAPR_MAX_LEVELS], e->stream);
    e->launches += 1;
    int rc;
    if ((rc = enqueue_resolve(e))) return rc;
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaEventRecord(e->ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (collect(e, graph, &r, true)) return 1; // miss: caller runs papr_ccdf_device on every shard
    for (int j = 0; j < r.nlevels; ++j) level_count[j] += r.level_count[j];
    return PAPR_OK;
}

// ------------------------------------------------------------------------------------------------
// stream-ordered shard stages (no host synchronisation until papr_shard_finish): the caller enqueues
// its collectives (NCCL) on papr_engine_stream() between them, directly on the engine's buffers.
// ------------------------------------------------------------------------------------------------
extern "C" int papr_engine_device_buffer(papr_engine *e, int which, void **ptr, uint64_t *bytes)
{
    if (!e || !ptr || !bytes) return PAPR_ERR_ARG;
    switch (which) {
    case PAPR_BUF_PRESAMPLE: *ptr = e->d_pre4; *bytes = 4 * sizeof(double); break;
    case PAPR_BUF_LOCAL_STATS: *ptr = &e->d_out->local; *bytes = sizeof(PaprDevStats); break;
    case PAPR_BUF_COUNTS: *ptr = e->d_out->counts; *bytes = sizeof(u64) * (PAPR_MAX_LEVELS + 1); break;
    default: return fail(e, PAPR_ERR_ARG, "unknown buffer id");
    }
    return PAPR_OK;
}

extern "C" int papr_shard_presample_async(papr_engine *e, const float *d_iq, uint64_t n, int graph)
{
    if (!e || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    begin_analysis(e);
    CU(cudaEventRecord(e->ev_begin, e->stream));
    papr_launch_presample(d_iq, std::min<u64>(n, kMaxLaunchSamples * 2 - PAPR_BATCH_SAMPLES),
                          presample_stride_for(e, n, graph ? 1 : 0), e->grid, e->d_pre_cta, e->stream);
    papr_launch_presample_reduce(e->d_pre_cta, e->grid, e->d_pre4, e->stream);
    e->launches += 2;
    CU(cudaGetLastError());
    return PAPR_OK;
}

extern "C" int papr_shard_scan_async(papr_engine *e, const float *d_iq, uint64_t n, uint64_t first, int graph, int fused)
{
    if (!e || (n && !d_iq)) return PAPR_ERR_ARG;
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    graph = graph ? 1 : 0;
    int rc;
    e->shard_mode = fused ? PAPR_MODE_FUSED : PAPR_MODE_TWO_PASS;
    if (!fused) {
        begin_analysis(e);
        CU(cudaEventRecord(e->ev_begin, e->stream));
    }
    if ((rc = enqueue_reset(e))) return rc;
    if (fused) {
        papr_launch_plan_pred(e->d_pre4, nullptr, 0, e->tables(graph), e->window_sigmas, e->predict_bias,
                              1 << (e->fine_bytes_log2 - 3 - PAPR_SH_MIN), &e->d_out->plan, e->d_fine_base, e->stream);
        papr_launch_zero_fine(&e->d_out->plan, e->d_fine, e->num_sms * 4, e->stream);
        e->launches += 2;
    }
    if ((rc = enqueue_scan(e, true, fused != 0, d_iq, n, first, true))) return rc;
    papr_launch_stats_finalize(e->d_work->wp, e->grid, n, &e->d_out->local, e->stream);
    e->launches += 1;
    CU(cudaGetLastError());
    return PAPR_OK;
}

extern "C" int papr_shard_counts_async(papr_engine *e, const void *d_all_stats, int nparts, int graph, int fused,
                                       const float *d_iq, uint64_t n)
{
    if (!e || !d_all_stats || nparts < 1) return PAPR_ERR_ARG;
    graph = graph ? 1 : 0;
    int rc;
    papr_launch_levels((const PaprDevStats *)d_all_stats, nparts, e->tables(graph), graph, &e->d_out->merged,
                       &e->d_out->lv, &e->d_out->counts[PAPR_MAX_LEVELS], e->stream);
    e->launches += 1;
    if (fused) {
        if ((rc = enqueue_resolve(e))) return rc;
    } else {
        CU(cudaMemsetAsync(e->d_work, 0, offsetof(DevWork, wp), e->stream)); // histogram scratch only
        if ((rc = enqueue_hist_exact(e, d_iq, n, true))) return rc;
    }
    return PAPR_OK;
}

extern "C" int papr_shard_finish(papr_engine *e, int graph, papr_result *out)
{
    if (!e || !out) return PAPR_ERR_ARG;
    graph = graph ? 1 : 0;
    int rc;
    reset_result(out);
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaEventRecord(e->ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    stats_to_host(e->h_out->o.merged, &out->stats);
    int redo = collect(e, graph, out, true) ? 1 : 0;
    finish_timing(e, out);
    out->mode_used = e->shard_mode;
    out->fused_miss = redo;
    return redo; // 1: run papr_shard_counts_async(..., fused = 0, ...) on every rank, reduce, finish again
}

// ------------------------------------------------------------------------------------------------
// sharded analysis with the exchanges fused into the kernels over peer memory (NVLink): one call per
// rank, the same six launches as the single-GPU fused analysis
// ------------------------------------------------------------------------------------------------
extern "C" int papr_xchg_export(papr_engine *e, void *handle64)
{
    if (!e || !handle64) return PAPR_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaSetDevice(e->device);
    if (!e->d_xchg) {
        CU(cudaMalloc(&e->d_xchg, sizeof(PaprXchg)));
        CU(cudaMemset(e->d_xchg, 0, sizeof(PaprXchg)));
        CU(cudaDeviceSynchronize()); // zeroed before any peer can learn the handle
    }
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, e->d_xchg));
    memcpy(handle64, &h, 64);
    return PAPR_OK;
}

extern "C" int papr_xchg_attach(papr_engine *e, int rank, int world, const void *handles)
{
    if (!e || !handles || world < 1 || world > PAPR_XCHG_MAX_RANKS || rank < 0 || rank >= world) return PAPR_ERR_ARG;
    if (!e->d_xchg) return fail(e, PAPR_ERR_ARG, "papr_xchg_export must run first");
    if (e->xchg_attached) return fail(e, PAPR_ERR_ARG, "already attached");
    cudaSetDevice(e->device);
    PaprPeers pp = {};
    pp.rank = rank;
    pp.world = world;
    for (int q = 0; q < world; ++q) {
        if (q == rank) { pp.win[q] = e->d_xchg; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)handles + 64 * q, 64);
        void *ptr = nullptr;
        cudaError_t er = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (er != cudaSuccess) {
            for (int k = 0; k < q; ++k)
                if (k != rank && pp.win[k]) cudaIpcCloseMemHandle(pp.win[k]);
            cudaGetLastError();
            return fail(e, PAPR_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(er));
        }
        pp.win[q] = (PaprXchg *)ptr;
    }
    pp.timeout_ns = (u64)(e->xchg_timeout_s * 1e9);
    CU(cudaMemset(e->d_xchg->abort, 0, sizeof(e->d_xchg->abort))); // a fresh attachment starts unpoisoned
    CU(cudaDeviceSynchronize());
    e->peers = pp;
    e->xchg_attached = true;
    return PAPR_OK;
}

// Unmap the other ranks' windows (the own one stays).  Callers detach on every rank and synchronise
// among themselves BEFORE any rank destroys its engine: freeing an exported window that a peer still
// has open is undefined.
extern "C" int papr_xchg_detach(papr_engine *e)
{
    if (!e) return PAPR_ERR_ARG;
    if (!e->xchg_attached) return PAPR_OK;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int q = 0; q < e->peers.world; ++q)
        if (q != e->peers.rank && e->peers.win[q]) cudaIpcCloseMemHandle(e->peers.win[q]);
    e->peers = PaprPeers{};
    e->xchg_attached = false;
    cudaGetLastError();
    return PAPR_OK;
}

extern "C" int papr_shard_analyze_p2p(papr_engine *e, const float *d_iq, uint64_t n, uint64_t first, int graph,
                                      papr_result *out)
{
    if (!e || !out || (n && !d_iq)) return PAPR_ERR_ARG;
    if (!e->xchg_attached) return fail(e, PAPR_ERR_ARG, "papr_xchg_attach must run first");
    if ((uintptr_t)d_iq & 15) return fail(e, PAPR_ERR_ARG, "device pointer must be 16-byte aligned");
    graph = graph ? 1 : 0;
    begin_analysis(e);
    reset_result(out);
    out->mode_used = PAPR_MODE_FUSED;
    e->shard_mode = PAPR_MODE_FUSED;
    int rc;
    CU(cudaEventRecord(e->ev_begin, e->stream));
    if ((rc = enqueue_reset(e))) return rc;
    // the sequential sum inside the sweep, if this shard qualifies (tunable p2p_chain = 0: never).  A rank that does not
    // still takes part in the chain exchange - it declines there, and then every rank reports the fall-back - so the
    // sequence of exchanges never depends on the shard sizes and the ranks need not agree on anything beforehand.
    const bool chained = e->p2p_chain != 0 && xt_applicable(e, n);
    if (e->p2p_chain > 0 && !chained) return fail(e, PAPR_ERR_ARG, "p2p_chain = 1 but this shard is too small for the chained sweep");
    e->xt_status = -1;
    papr_launch_presample(d_iq, std::min<u64>(n, kMaxLaunchSamples * 2 - PAPR_BATCH_SAMPLES),
                          presample_stride_for(e, n, graph), e->grid, e->d_pre_cta, e->stream);
    papr_launch_plan_pred_x(e->d_pre_cta, e->grid, e->tables(graph), e->window_sigmas, e->predict_bias,
                            1 << (e->fine_bytes_log2 - 3 - PAPR_SH_MIN), &e->d_out->plan, e->d_fine_base, e->peers,
                            ++e->xseq[XK_PRE], e->stream, chained ? XT_NCELLS : PAPR_NCELLS_MAX);
    papr_launch_zero_fine(&e->d_out->plan, e->d_fine, e->num_sms * 4, e->stream);
    e->launches += 3;
    const unsigned ntiles = (unsigned)((n + XT_TILE_SAMPLES - 1) / XT_TILE_SAMPLES);
    // every rank's pass-1 state -> merged statistics and the levels of the merged FIXED-ORDER sums ...
    PaprFinalizeXArgs fx;
    fx.wp = e->d_work->wp; fx.nctas = e->grid; fx.n = n; fx.local = &e->d_out->local; fx.tb = e->tables(graph); fx.graph = graph;
    fx.merged = &e->d_out->merged; fx.lv = &e->d_out->lv; fx.status_word = &e->d_out->counts[PAPR_MAX_LEVELS];
    fx.plan = &e->d_out->plan; fx.pp = e->peers; fx.seq = ++e->xseq[XK_STATS]; fx.bias = e->epilogue_bias;
    if (chained) { // ... as one extra CTA of the kernel that composes the tile runs
        if ((rc = enqueue_scan_tma(e, d_iq, n, first, true))) return rc;
        const unsigned nsuper = (ntiles + XT_SUPER_TILES - 1) / XT_SUPER_TILES, nhyper = (nsuper + XT_HYPER_SUPERS - 1) / XT_HYPER_SUPERS;
        papr_launch_xt_compose(e->d_xt_run, e->d_xt_code, ntiles, e->d_xt_multi_tile, e->d_xt_super, e->d_xt_hyper,
                               (int)std::min<unsigned>(nhyper, (unsigned)e->num_sms * 2), e->stream, nullptr, &fx);
    } else {
        if ((rc = enqueue_scan(e, true, true, d_iq, n, first, true))) return rc;
        papr_launch_finalize_levels_x(fx, e->stream);
    }
    e->launches += 1;
    // ... then the chain (with its exchange) side by side with this shard's counts against those levels
    papr_launch_xt_epilogue_x(e->d_xt_hyper, e->d_xt_super, e->d_xt_run, e->d_xt_code, e->d_xt_multi, e->d_xt_multi_tile, ntiles,
                              d_iq, n, e->d_xt_chain, &e->d_out->plan, e->peers, ++e->xseq[XK_CHAIN], epilogue_args(e, n, graph),
                              e->num_sms, e->stream, chained ? 0 : 1, ++e->xseq[XK_COUNTS]); // (counts exchange in its tail)
    e->launches += 1;
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaEventRecord(e->ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    if (e->h_out->o.plan.pad) return fail(e, PAPR_ERR_INTERNAL, "peer exchange timed out (a rank did not take part)");
    stats_to_host(e->h_out->o.merged, &out->stats);
    { // identical on every rank: all of them walked the same lists (a rank that declined makes everyone fall back)
        e->xt_status = e->h_out->o.chain[0];
        e->xt_why = e->h_out->o.chain[1];
        if (e->xt_status == XT_OK) out->stats.sum = e->h_out->o.chain_exact; // replaces the merged fixed-order sums
        out->sum_path = e->xt_status == XT_OK ? 1u : e->xt_status == XT_FALLBACK ? (2u | ((unsigned)e->xt_why << 8)) : 0u;
        if (e->exact_sum == 0) out->sum_path = 0; // any fixed order was asked for
    }
    // merged stats, the chained sum and the summed status word are identical on every rank, so every rank takes
    // the same branches here
    int redo = collect(e, graph, out, true);
    if (redo & 1) {
        // only the levels moved (they were derived from the fixed-order sums while the chain was running): every
        // rank counts again against the levels of the chained sum - the cells and fine tables are still in place
        if ((rc = upload_levels(e, out->level, out->nlevels, out->stats.peak))) return rc;
        if ((rc = enqueue_resolve(e))) return rc;
        papr_launch_counts_x(e->d_out->counts, &e->d_out->lv, &e->d_out->plan, e->peers, ++e->xseq[XK_COUNTS], e->stream);
        e->launches += 1;
        if ((rc = enqueue_fetch(e))) return rc;
        CU(cudaEventRecord(e->ev_end, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        if (e->h_out->o.plan.pad) return fail(e, PAPR_ERR_INTERNAL, "peer exchange timed out (a rank did not take part)");
        redo = e->h_out->o.counts[PAPR_MAX_LEVELS] != 0 ? 2 : 0;
        if (!redo) {
            out->fused_miss = 2;
            for (int j = 0; j < out->nlevels; ++j) out->level_count[j] = (int64_t)e->h_out->o.counts[j];
        }
    }
    // thresholds outside the predicted windows somewhere -> exact pass everywhere
    if (redo) {
        out->fused_miss = 1;
        if ((rc = upload_levels(e, out->level, out->nlevels, out->stats.peak))) return rc;
        CU(cudaMemsetAsync(e->d_work, 0, offsetof(DevWork, wp), e->stream)); // histogram scratch only
        if ((rc = enqueue_hist_exact(e, d_iq, n, true))) return rc;
        papr_launch_counts_x(e->d_out->counts, &e->d_out->lv, &e->d_out->plan, e->peers, ++e->xseq[XK_COUNTS], e->stream);
        e->launches += 1;
        if ((rc = enqueue_fetch(e))) return rc;
        CU(cudaEventRecord(e->ev_end, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        if (e->h_out->o.plan.pad) return fail(e, PAPR_ERR_INTERNAL, "peer exchange timed out (a rank did not take part)");
        if (e->h_out->o.counts[PAPR_MAX_LEVELS] != 0) return fail(e, PAPR_ERR_INTERNAL, "exact CCDF pass reported a miss");
        for (int j = 0; j < out->nlevels; ++j) out->level_count[j] = (int64_t)e->h_out->o.counts[j];
    }
    finish_timing(e, out);
    return PAPR_OK;
}

// tone-reservation PAPR reduction of nsym time-domain OFDM symbols of fft_size samples each, in place (papr_tr.cu)
extern "C" int papr_tr_reduce_device(papr_engine *e, float *d_symbols, int nsym, int fft_size, const float *d_kernel,
                                     const int *d_tones, int ntones, float vclip, int iterations, float amax,
                                     float *d_tone_values, int *d_iterations)
{
    if (!e || !d_symbols || !d_kernel || !d_tones || !d_tone_values || nsym < 0 || fft_size < 2 || ntones < 1 || iterations < 0)
        return PAPR_ERR_ARG;
    if (!(vclip > 0.f) || !(amax >= 0.f)) return fail(e, PAPR_ERR_ARG, "vclip must be positive and amax non-negative");
    cudaSetDevice(e->device);
    if (nsym == 0) return PAPR_OK;
    papr_launch_tr(d_symbols, nsym, fft_size, d_kernel, d_tones, ntones, vclip, iterations, amax, d_tone_values, d_iterations,
                   std::min(nsym, e->num_sms * 2), e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return PAPR_OK;
}

extern "C" int papr_siggen_device(papr_engine *e, float *d_iq, uint64_t first, uint64_t n, uint64_t seed)
{
    if (!e || (n && !d_iq)) return PAPR_ERR_ARG;
    cudaSetDevice(e->device);
    papr_launch_siggen(d_iq, first, n, seed, e->num_sms * 16, e->stream);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(e->stream));
    return PAPR_OK;
}

// ------------------------------------------------------------------------------------------------
// host-side captures: chunked H2D (direct from pinned sources; through a pinned staging ring filled by
// helper threads with memcpy / pread for pageable images and files), the statistics pass - and, chunk
// by chunk, the exact sequential-sum emulation - running on each chunk as it lands, so that only the
// CCDF pass is left when the last byte has crossed PCIe.  The shard stays resident in HBM for that
// pass; a capture larger than the resident budget is streamed a second time instead, through a ring
// of device chunk buffers (the reference reads its file twice as well, papr.c:142).
// ------------------------------------------------------------------------------------------------
static int ensure_device_buffer(papr_engine *e, size_t bytes)
{
    bytes = (bytes + 4095) & ~(size_t)4095;
    if (bytes <= e->d_buf_bytes) return PAPR_OK;
    if (e->d_buf) { cudaFree(e->d_buf); e->d_buf = nullptr; e->d_buf_bytes = 0; }
    CU(cudaMalloc(&e->d_buf, bytes));
    e->d_buf_bytes = bytes;
    return PAPR_OK;
}

static int ensure_ring(papr_engine *e)
{
    if (e->d_ring && e->d_ring_chunk == e->chunk_bytes) return PAPR_OK;
    if (e->d_ring) { cudaFree(e->d_ring); e->d_ring = nullptr; }
    if (e->d_buf) { cudaFree(e->d_buf); e->d_buf = nullptr; e->d_buf_bytes = 0; } // make room
    CU(cudaMalloc(&e->d_ring, (size_t)kRing * e->chunk_bytes + 16));
    e->d_ring_chunk = e->chunk_bytes;
    return PAPR_OK;
}

static int staging_threads_of(const papr_engine *e)
{
    int n = e->staging_threads;
    if (n < 0) n = std::min(16, std::max(2, ((int)std::thread::hardware_concurrency() - 2) / std::max(1, e->host_share)));
    return std::min(kMaxStages - 2, std::max(1, n));
}

static size_t piece_bytes_of(const papr_engine *e)
{
    size_t p = std::min(kPieceBytes, e->chunk_bytes);
    while (e->chunk_bytes % p) p >>= 1; // chunk_bytes is a multiple of 256 KiB, so this stops there at the latest
    return p;
}

// one pinned slot (one staging piece) per helper thread plus two, so that a copy can be in flight
// while every helper reads
static int ensure_staging(papr_engine *e)
{
    const int slots = staging_threads_of(e) + 2;
    const size_t piece = piece_bytes_of(e);
    if (e->h_stage_bytes != piece || e->h_stage_slots != slots) {
        if (e->h_stage[0]) cudaFreeHost(e->h_stage[0]); // one allocation, carved into the slots
        for (auto &p : e->h_stage) p = nullptr;
        e->h_stage_bytes = 0;
        void *block = nullptr;
        CU(cudaHostAlloc(&block, piece * (size_t)slots, cudaHostAllocDefault));
        for (int i = 0; i < slots; ++i) e->h_stage[i] = (char *)block + piece * (size_t)i;
        e->h_stage_bytes = piece;
        e->h_stage_slots = slots;
    }
    return PAPR_OK;
}

// does a shard of `bytes` stay resident?  (tunable "max_resident_bytes"; default: what the device has
// free right now - counting the buffer this engine already holds - minus a 1 GiB reserve)
static bool fits_resident(papr_engine *e, u64 bytes)
{
    if (e->max_resident_bytes) return bytes <= e->max_resident_bytes;
    if (bytes <= e->d_buf_bytes) return true;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return true; }
    const u64 avail = (u64)free_b + e->d_buf_bytes;
    return bytes + (1ull << 30) <= avail;
}

// geometry of one host-side shard
struct StreamGeom {
    u64 npairs = 0, n = 0;     // complete pairs in the source; samples incl. the lone-I tail sample
    bool tail = false;
    float tail_pair[2] = {0.f, 0.f};
    u64 chunk_samples = 0, first = 0;
};

// The lone trailing I of a capture with an odd number of floats is paired with a stale Q
// (papr_host_stale_q explains which); same rule, reading through the source
static float stale_q_of(const HostSource &src)
{
    if (src.img) return papr_host_stale_q(src.img, src.bytes);
    const u64 chunk = 16384, nfloats = src.bytes / 4, last = nfloats - 1; // CHUNK_SIZE, papr.c:30
    const u64 slot = last % chunk + 1, chunk_start = last - last % chunk;
    unsigned char q[4] = {0, 0, 0, 0};
    float f;
    if (chunk_start >= chunk) src.read(q, 4 * (chunk_start - chunk + slot), 4);
    if (src.bytes % 4) src.read(q, 4 * nfloats, src.bytes % 4);
    memcpy(&f, q, 4);
    return f;
}

static int make_geom(papr_engine *e, const HostSource &src, u64 first, const float *tail_override, StreamGeom *g)
{
    const u64 nfloats = src.bytes / 4;
    g->npairs = nfloats / 2;
    g->tail = tail_override ? true : (nfloats & 1) != 0;
    g->n = g->npairs + (g->tail ? 1 : 0); // papr.c:102: a lone trailing I still counts as a sample
    g->first = first;
    g->chunk_samples = e->chunk_bytes / 8;
    if (tail_override) {
        g->tail_pair[0] = tail_override[0];
        g->tail_pair[1] = tail_override[1];
    } else if (g->tail) {
        if (!src.read(&g->tail_pair[0], 4 * (nfloats - 1), 4)) return fail(e, PAPR_ERR_IO, "read failed");
        g->tail_pair[1] = stale_q_of(src);
    }
    return PAPR_OK;
}

// One PCIe pass: chunk c = samples [off, off+m) lands at d_chunk, then work(d_chunk, off, m, c, last)
// enqueues on e->stream.  Re-streaming: work must call release(c') (recorded on e->stream) once
// everything that reads chunk c' is enqueued; its ring slot is refilled only after that.
typedef std::function<void(u64 c)> ChunkRelease;
typedef std::function<int(const float *d_chunk, u64 off, u64 m, u64 c, bool last, const ChunkRelease &release)> ChunkWork;

static int stream_chunks(papr_engine *e, const HostSource &src, const StreamGeom &g, bool resident, const ChunkWork &work)
{
    int rc;
    const ChunkRelease release = [&](u64 c) {
        if (!resident) cudaEventRecord(e->ring_free[c % kRing], e->stream);
    };
    std::unique_ptr<ChunkFeeder> feeder; // helper threads stop and join when this goes out of scope
    const u64 piece = piece_bytes_of(e);
    if (!src.pinned && g.npairs)
        feeder.reset(new ChunkFeeder(e->device, src, (g.npairs * 8 + piece - 1) / piece, piece, g.npairs * 8,
                                     staging_threads_of(e), e->h_stage, e->stage_done, e->h_stage_slots));
    u64 c = 0;
    for (u64 off = 0; off < g.n; off += g.chunk_samples, ++c) {
        const u64 m = std::min(g.chunk_samples, g.n - off);               // samples of this chunk (incl. the tail sample)
        const u64 mp = std::min(m, g.npairs > off ? g.npairs - off : 0);  // complete pairs to copy from the source
        float *dst = resident ? e->d_buf + 2 * off : e->d_ring + (c % kRing) * (e->d_ring_chunk / 4);
        if (!resident) CU(cudaStreamWaitEvent(e->copy_stream, e->ring_free[c % kRing], 0)); // previous tenant consumed?
        if (mp && src.pinned) {
            CU(cudaMemcpyAsync(dst, src.img + off * 8, mp * 8, cudaMemcpyHostToDevice, e->copy_stream));
        } else if (mp) { // piece by piece through the staging slots (chunks are whole pieces)
            for (u64 b = 0; b < mp * 8; b += piece) {
                const u64 p = (off * 8 + b) / piece;
                const void *from = feeder->wait_filled(p);
                if (!from) return fail(e, PAPR_ERR_IO, "read failed while streaming the capture");
                CU(cudaMemcpyAsync((char *)dst + b, from, std::min(piece, mp * 8 - b), cudaMemcpyHostToDevice, e->copy_stream));
                CU(cudaEventRecord(e->stage_done[p % e->h_stage_slots], e->copy_stream));
                feeder->issued(p);
            }
        }
        e->h2d += mp * 8;
        if (g.tail && off + m == g.n) {
            CU(cudaMemcpyAsync(dst + 2 * (g.npairs - off), g.tail_pair, 8, cudaMemcpyHostToDevice, e->copy_stream));
            e->h2d += 8;
        }
        CU(cudaEventRecord(e->chunk_ready, e->copy_stream));
        CU(cudaStreamWaitEvent(e->stream, e->chunk_ready, 0));
        if ((rc = work(dst, off, m, c, off + m == g.n, release))) return rc;
    }
    return PAPR_OK;
}

// brings a tile's samples to the host for a literal replay (seq_chain): D2H when the shard is
// resident, otherwise a read from the source (+ the synthesised tail pair)
static TileFetch tile_fetcher(papr_engine *e, const HostSource &src, const StreamGeom &g, bool resident)
{
    return [e, &src, &g, resident](u64 first, u64 cnt, float *dst) -> int {
        if (resident) {
            CU(cudaMemcpyAsync(dst, e->d_buf + 2 * first, cnt * 8, cudaMemcpyDeviceToHost, e->stream));
            CU(cudaStreamSynchronize(e->stream));
            e->d2h += cnt * 8;
            return PAPR_OK;
        }
        const u64 pairs = std::min(cnt, g.npairs > first ? g.npairs - first : 0);
        if (pairs && !src.read(dst, first * 8, pairs * 8)) return fail(e, PAPR_ERR_IO, "read failed while replaying a tile");
        if (pairs < cnt) { dst[2 * pairs] = g.tail_pair[0]; dst[2 * pairs + 1] = g.tail_pair[1]; }
        return PAPR_OK;
    };
}

// Pass 1 over a host-side shard: statistics per chunk as it lands; with inline_exact also the exact
// sequential sum (tile sums per chunk, tiles placed in binades kSeqLag chunks later, tile runs on the
// GPU while later chunks are still crossing PCIe; the host chains them at the end).  On return the
// local state is finalized in d_out->local (sum already exact if *exact_done) - nothing synchronised
// unless inline_exact.
enum SeqInline { SEQ_NONE = 0, SEQ_SUMS_ONLY = 1, SEQ_FULL = 2 };

// who delivers the chunks: a seekable source streamed by stream_chunks, or a pipe read once (stream_fd_chunks,
// which also fills in the geometry - the length of a pipe is known only at its end)
typedef std::function<int(const ChunkWork &work)> ChunkDriver;

static int host_pass1_driver(papr_engine *e, StreamGeom &g, u64 max_tiles, bool resident, SeqInline seq, bool *exact_done,
                             const ChunkDriver &drive, const TileFetch &fetch)
{
    int rc;
    *exact_done = false;
    const bool inline_exact = seq == SEQ_FULL;
    if (seq != SEQ_NONE && (rc = ensure_seq_buffers(e, std::max<u64>(max_tiles, 1)))) return rc;
    if ((rc = enqueue_reset(e))) return rc;

    struct Pending { const float *d; u64 off, m, c; };
    std::vector<Pending> pend;
    size_t pend_head = 0;
    bool seq_ok = inline_exact;
    double pre = 0.0;
    e->seq_dirty = 0;
    auto process = [&](const Pending &ch, const ChunkRelease &release) -> int {
        const u64 tile0 = ch.off / PAPR_SEQ_TILE, nt = (ch.m + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
        CU(cudaEventSynchronize(e->seq_ev[ch.c % kSeqEvents])); // this chunk's tile sums are on the host
        if (seq_ok && seq_classify(e->h_tile_sum + tile0, nt, &pre, e->h_tile_code + tile0)) seq_ok = false;
        if (seq_ok) {
            CU(cudaMemcpyAsync(e->d_tile_code + tile0, e->h_tile_code + tile0, nt * sizeof(short), cudaMemcpyHostToDevice,
                               e->stream));
            papr_launch_seqsum(ch.d, ch.m, e->d_tile_code + tile0, e->d_tile_run + 2 * tile0,
                               (int)std::min<u64>((u64)e->num_sms * 8, nt), e->stream);
            e->launches += 1;
            e->h2d += nt * 2;
        }
        release(ch.c);
        return PAPR_OK;
    };
    rc = drive([&](const float *d, u64 off, u64 m, u64 c, bool last, const ChunkRelease &release) -> int {
        int r = enqueue_scan(e, true, false, d, m, g.first + off, false);
        if (r) return r;
        if (seq == SEQ_NONE) { release(c); return PAPR_OK; }
        const u64 tile0 = off / PAPR_SEQ_TILE, nt = (m + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
        papr_launch_tilesum(d, m, e->d_tile_sum + tile0, (int)std::min<u64>((u64)e->num_sms * 2, nt), e->stream);
        e->launches += 1;
        CU(cudaMemcpyAsync(e->h_tile_sum + tile0, e->d_tile_sum + tile0, nt * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        e->d2h += nt * 8;
        if (seq == SEQ_SUMS_ONLY) { release(c); return PAPR_OK; } // the caller classifies and chains across shards
        CU(cudaEventRecord(e->seq_ev[c % kSeqEvents], e->stream));
        pend.push_back({d, off, m, c});
        while (pend.size() - pend_head > (size_t)(last ? 0 : (resident ? kSeqLagResident : kSeqLagRing)))
            if ((r = process(pend[pend_head++], release))) return r;
        return PAPR_OK;
    });
    if (rc) return rc;
    const u64 ntiles = (g.n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
    papr_launch_stats_finalize(e->d_work->wp, e->grid, g.n, &e->d_out->local, e->stream);
    e->launches += 1;
    CU(cudaGetLastError());
    if (!seq_ok || ntiles == 0) return PAPR_OK;
    // the runs come back, the host chains them in file order, the exact sum replaces the tree sum
    CU(cudaMemcpyAsync(e->h_tile_run, e->d_tile_run, ntiles * 2 * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->d2h += ntiles * 16;
    double s = 0.0;
    rc = seq_chain(e, g.n, &s, fetch);
    if (rc) return rc;
    e->h_out->pre4[0] = s;
    CU(cudaMemcpyAsync(&e->d_out->local.sum, &e->h_out->pre4[0], sizeof(double), cudaMemcpyHostToDevice, e->stream));
    *exact_done = true;
    return PAPR_OK;
}

static int host_pass1(papr_engine *e, const HostSource &src, const StreamGeom &g, bool resident, SeqInline seq,
                      bool *exact_done)
{
    StreamGeom gg = g;
    return host_pass1_driver(e, gg, (g.n + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE, resident, seq, exact_done,
                             [&](const ChunkWork &work) { return stream_chunks(e, src, g, resident, work); },
                             tile_fetcher(e, src, g, resident));
}

// CCDF pass of the re-streaming mode: the capture crosses PCIe a second time, chunk by chunk
static int host_hist_streamed(papr_engine *e, const HostSource &src, const StreamGeom &g)
{
    int rc;
    if ((rc = hist_begin(e))) return rc;
    rc = stream_chunks(e, src, g, false, [&](const float *d, u64, u64 m, u64 c, bool, const ChunkRelease &release) -> int {
        int r = hist_chunk(e, d, m, false);
        release(c);
        return r;
    });
    if (rc) return rc;
    return hist_end(e);
}

// NaN sign (see fix_nan_sign) when the shard is not resident: one more pass, only ever taken for
// captures larger than HBM that contain a NaN
static int fix_nan_sign_streamed(papr_engine *e, const HostSource &src, const StreamGeom &g, papr_stats *st)
{
    if (!std::isnan(st->sum) || g.n == 0) return PAPR_OK;
    CU(cudaMemsetAsync(e->d_nan_idx, 0xff, sizeof(u64), e->stream));
    int rc = stream_chunks(e, src, g, false, [&](const float *d, u64 off, u64 m, u64 c, bool, const ChunkRelease &release) -> int {
        papr_launch_find_nan(d, m, g.first + off, e->d_nan_idx, e->num_sms * 8, e->stream);
        e->launches += 1;
        release(c);
        return PAPR_OK;
    });
    if (rc) return rc;
    CU(cudaMemcpyAsync(&e->h_out->nan_idx, e->d_nan_idx, sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    const u64 k = e->h_out->nan_idx;
    if (k == ~0ull) return PAPR_OK;
    float pair[2];
    rc = tile_fetcher(e, src, g, false)(k - g.first, 1, pair);
    if (rc) return rc;
    const bool neg = std::isnan(pair[1]) ? std::signbit(pair[1]) : std::signbit(pair[0]);
    st->sum = std::copysign(std::fabs(st->sum), neg ? -1.0 : 1.0);
    return PAPR_OK;
}

static bool is_pinned_host(const void *p, u64 bytes)
{
    cudaPointerAttributes attr;
    const bool pinned = bytes && cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

static HostSource memory_source(const void *image, u64 bytes)
{
    static const unsigned char none = 0;
    HostSource src;
    src.img = image ? (const unsigned char *)image : &none; // empty capture
    src.bytes = bytes;
    src.pinned = is_pinned_host(image, bytes);
    return src;
}

static int analyze_source(papr_engine *e, const HostSource &src, int graph, papr_result *out)
{
    graph = graph ? 1 : 0;
    begin_analysis(e);
    reset_result(out);
    out->mode_used = PAPR_MODE_TWO_PASS;
    int rc;
    StreamGeom g;
    if ((rc = make_geom(e, src, 0, nullptr, &g))) return rc;
    const bool resident = fits_resident(e, g.n * 8 + 16);
    e->streamed = !resident;
    const bool trace = getenv("PAPR_B200_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    auto mark = [&](const char *what) {
        if (trace)
            fprintf(stderr, "papr_b200 trace: %-34s %9.3f ms\n", what,
                    std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };
    if (resident) rc = ensure_device_buffer(e, g.n * 8 + 16);
    else rc = ensure_ring(e);
    if (rc) return rc;
    mark(resident ? "device buffer (resident shard)" : "device buffers (re-streaming ring)");
    if (!src.pinned && g.npairs && (rc = ensure_staging(e))) return rc;
    mark("pinned staging slots");
    CU(cudaEventRecord(e->ev_begin, e->stream));
    bool exact_done = false;
    if ((rc = host_pass1(e, src, g, resident, e->exact_sum != 0 ? SEQ_FULL : SEQ_NONE, &exact_done))) return rc;
    out->sum_path = exact_done ? 3u : 0u;
    mark("H2D + pass 1 (+ exact sum)");
    // local state -> merged state, avg, L, levels on the device
    papr_launch_levels(&e->d_out->local, 1, e->tables(graph), graph, &e->d_out->merged, &e->d_out->lv,
                       &e->d_out->counts[PAPR_MAX_LEVELS], e->stream);
    e->launches += 1;
    if (resident) rc = enqueue_hist_exact(e, e->d_buf, g.n, true);
    else rc = host_hist_streamed(e, src, g);
    if (rc) return rc;
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaEventRecord(e->ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    mark("CCDF pass + fetch");
    stats_to_host(e->h_out->o.merged, &out->stats);
    if (resident) rc = fix_nan_sign(e, e->d_buf, g.n, 0, &out->stats);
    else rc = fix_nan_sign_streamed(e, src, g, &out->stats);
    if (rc) return rc;
    if (collect(e, graph, out, true)) { // device level table != host libm's: redo with the host's
        if (resident) {
            rc = run_exact_ccdf(e, e->d_buf, g.n, out->level, out->nlevels, out->stats.peak, out->level_count, false);
        } else {
            if ((rc = upload_levels(e, out->level, out->nlevels, out->stats.peak))) return rc;
            if ((rc = enqueue_reset(e))) return rc;
            if ((rc = host_hist_streamed(e, src, g))) return rc;
            if ((rc = enqueue_fetch(e))) return rc;
            CU(cudaStreamSynchronize(e->stream));
            if (e->h_out->o.counts[PAPR_MAX_LEVELS] != 0) return fail(e, PAPR_ERR_INTERNAL, "exact CCDF pass reported a miss");
            for (int j = 0; j < out->nlevels; ++j) out->level_count[j] = (int64_t)e->h_out->o.counts[j];
        }
        if (rc) return rc;
        CU(cudaEventRecord(e->ev_end, e->stream));
        CU(cudaStreamSynchronize(e->stream));
    }
    finish_timing(e, out);
    return PAPR_OK;
}

extern "C" int papr_analyze_host(papr_engine *e, const void *image, uint64_t bytes, int graph, papr_result *out)
{
    if (!e || !out || (bytes && !image)) return PAPR_ERR_ARG;
    cudaSetDevice(e->device);
    return analyze_source(e, memory_source(image, bytes), graph, out);
}

// pass 1 of a host-side shard that must stay resident (sharded callers run their own exchange
// before the CCDF pass); nothing synchronised on return
static int host_stream_stats(papr_engine *e, const HostSource &src, u64 first, u64 *n_out,
                             const float *tail_override = nullptr)
{
    StreamGeom g;
    int rc;
    if ((rc = make_geom(e, src, first, tail_override, &g))) return rc;
    *n_out = g.n;
    if (!fits_resident(e, g.n * 8 + 16))
        return fail(e, PAPR_ERR_ARG, "shard exceeds the resident budget of this GPU: use more shards");
    if ((rc = ensure_device_buffer(e, g.n * 8 + 16))) return rc;
    if (!src.pinned && g.npairs && (rc = ensure_staging(e))) return rc;
    // tile sums of the sequential-sum emulation ride along under the H2D shadow; the caller places and
    // chains the tiles across shards afterwards (papr_seqsum_* / papr_multi), once it has synchronised
    bool exact_done = false;
    const bool sums = e->exact_sum != 0 && g.n > 0;
    if ((rc = host_pass1(e, src, g, true, sums ? SEQ_SUMS_ONLY : SEQ_NONE, &exact_done))) return rc;
    if (sums) { e->seq_sums_ptr = e->d_buf; e->seq_sums_n = g.n; }
    return PAPR_OK;
}

// Sharded callers with host-resident shards: pass 1 while streaming in; the shard stays resident at
// *d_iq_out (engine-owned, valid until the next host-path call) for papr_ccdf_device.
extern "C" int papr_stats_host(papr_engine *e, const void *image, uint64_t bytes, uint64_t first, papr_stats *out,
                               const float **d_iq_out, uint64_t *nsamples_out)
{
    if (!e || !out || (bytes && !image)) return PAPR_ERR_ARG;
    begin_analysis(e);
    int rc;
    u64 n = 0;
    cudaSetDevice(e->device);
    if ((rc = host_stream_stats(e, memory_source(image, bytes), first, &n))) return rc;
    CU(cudaMemcpyAsync(&e->h_out->o.local, &e->d_out->local, sizeof(PaprDevStats), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    stats_to_host(e->h_out->o.local, out);
    if (d_iq_out) *d_iq_out = e->d_buf;
    if (nsamples_out) *nsamples_out = n;
    return fix_nan_sign(e, e->d_buf, n, first, out);
}

// ------------------------------------------------------------------------------------------------
// non-seekable inputs (FIFOs, pipes, character devices): the reference cannot read them at all - it rewinds
// its file (papr.c:142).  Here the stream is read ONCE: every 4 MiB piece goes from read(2) into a pinned slot
// and straight on to the GPU, pass 1 and the sequential-sum stages run on each 64 MiB chunk while the next is
// still arriving, and the chunks stay resident in HBM for the CCDF pass.  Limit: the stream must fit the
// GPU's free memory (less 1 GiB).
// ------------------------------------------------------------------------------------------------
struct StreamChunks {
    std::vector<float *> dev; // one device buffer per chunk of chunk_samples (the last may hold fewer)
    u64 chunk_samples = 0;
    ~StreamChunks() { for (auto p : dev) cudaFree(p); }
};

static int stream_fd_chunks(papr_engine *e, int fd, StreamGeom *g, StreamChunks *sc, u64 max_bytes, const ChunkWork &work)
{
    int rc;
    const size_t piece = piece_bytes_of(e);
    const u64 chunk_bytes = e->chunk_bytes;
    const ChunkRelease release = [](u64) {};
    sc->chunk_samples = chunk_bytes / 8;
    g->chunk_samples = sc->chunk_samples;
    u64 total = 0;      // bytes read so far
    u64 fill = 0;       // bytes in the current (last) chunk
    u64 pending_c = 0;  // chunks handed to `work` so far
    bool have_full = false; // the last chunk is full and not yet handed on (is it the final one? only the next read tells)
    std::vector<unsigned char> recent; // the last <= 256 KiB of the stream (the stale Q of a lone trailing I lives there)
    u64 recent_base = 0;
    int slot = 0;
    auto dispatch = [&](bool last) -> int {
        const u64 c = pending_c++;
        const u64 off = c * sc->chunk_samples;
        const u64 m = last ? g->n - off : sc->chunk_samples;
        CU(cudaEventRecord(e->chunk_ready, e->copy_stream));
        CU(cudaStreamWaitEvent(e->stream, e->chunk_ready, 0));
        return work(sc->dev[c], off, m, c, last, release);
    };
    for (;;) {
        CU(cudaEventSynchronize(e->stage_done[slot])); // the slot's previous tenant has crossed PCIe
        unsigned char *buf = (unsigned char *)e->h_stage[slot];
        size_t got = 0;
        while (got < piece) {
            ssize_t k = ::read(fd, buf + got, piece - got);
            if (k < 0 && errno == EINTR) continue;
            if (k < 0) return fail(e, PAPR_ERR_IO, "read failed on the input stream");
            if (k == 0) break;
            got += (size_t)k;
        }
        if (got == 0) break;
        if (total + got > max_bytes)
            return fail(e, PAPR_ERR_ARG, "the input stream does not fit this GPU's memory; write it to a file first");
        for (size_t done = 0; done < got;) { // a piece may straddle two chunks
            if (have_full) { have_full = false; if ((rc = dispatch(false))) return rc; fill = 0; }
            if (fill == 0) {
                float *p = nullptr;
                CU(cudaMalloc(&p, chunk_bytes + 16));
                sc->dev.push_back(p);
            }
            const size_t take = (size_t)std::min<u64>(got - done, chunk_bytes - fill);
            CU(cudaMemcpyAsync((char *)sc->dev.back() + fill, buf + done, take, cudaMemcpyHostToDevice, e->copy_stream));
            fill += take; done += take;
            if (fill == chunk_bytes) have_full = true;
        }
        CU(cudaEventRecord(e->stage_done[slot], e->copy_stream));
        e->h2d += got;
        // remember the tail of the stream
        if (got >= (256u << 10)) { recent.assign(buf + got - (256u << 10), buf + got); recent_base = total + got - (256u << 10); }
        else {
            recent.insert(recent.end(), buf, buf + got);
            if (recent.size() > (512u << 10)) { const size_t cut = recent.size() - (256u << 10); recent.erase(recent.begin(), recent.begin() + cut); recent_base += cut; }
        }
        total += got;
        slot = (slot + 1) % e->h_stage_slots;
        if (got < piece) break; // end of stream
    }
    // geometry, now that the length is known (papr.c:101-103: a lone trailing I still counts as a sample)
    const u64 nfloats = total / 4;
    g->npairs = nfloats / 2;
    g->tail = (nfloats & 1) != 0;
    g->n = g->npairs + (g->tail ? 1 : 0);
    if (g->tail) {
        auto byte_at = [&](u64 off) -> unsigned char { return off >= recent_base && off - recent_base < recent.size() ? recent[off - recent_base] : 0; };
        const u64 chunk = 16384, last = nfloats - 1, slot_f = last % chunk + 1, chunk_start = last - last % chunk; // papr_host_stale_q
        unsigned char q[4] = {0, 0, 0, 0}, ib[4];
        for (int k = 0; k < 4; ++k) ib[k] = byte_at(4 * last + k);
        if (chunk_start >= chunk) for (int k = 0; k < 4; ++k) q[k] = byte_at(4 * (chunk_start - chunk + slot_f) + k);
        for (u64 k = 0; k < total % 4; ++k) q[k] = byte_at(4 * nfloats + k);
        memcpy(&g->tail_pair[0], ib, 4);
        memcpy(&g->tail_pair[1], q, 4);
        // the tail sample sits right after the last complete pair: possibly the first sample of a fresh chunk
        const u64 cidx = g->npairs / sc->chunk_samples;
        if (cidx == sc->dev.size()) {
            if (have_full) { have_full = false; if ((rc = dispatch(false))) return rc; }
            float *p = nullptr;
            CU(cudaMalloc(&p, 64));
            sc->dev.push_back(p);
        }
        CU(cudaMemcpyAsync(sc->dev[cidx] + 2 * (g->npairs - cidx * sc->chunk_samples), g->tail_pair, 8, cudaMemcpyHostToDevice, e->copy_stream));
        e->h2d += 8;
    }
    if (g->n > pending_c * sc->chunk_samples) return dispatch(true);
    return PAPR_OK;
}

static int analyze_stream(papr_engine *e, int fd, int graph, papr_result *out)
{
    graph = graph ? 1 : 0;
    begin_analysis(e);
    reset_result(out);
    out->mode_used = PAPR_MODE_TWO_PASS;
    int rc;
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const u64 budget = e->max_resident_bytes ? e->max_resident_bytes : (free_b > (1ull << 30) ? free_b - (1ull << 30) : 0);
    if ((rc = ensure_staging(e))) return rc;
    StreamGeom g;
    StreamChunks sc;
    const bool exact = e->exact_sum != 0;
    CU(cudaEventRecord(e->ev_begin, e->stream));
    bool exact_done = false;
    const TileFetch fetch = [&](u64 first, u64 cnt, float *dst) -> int { // a tile never straddles chunks (chunks are whole tiles)
        const u64 c = first / sc.chunk_samples;
        CU(cudaMemcpyAsync(dst, sc.dev[c] + 2 * (first - c * sc.chunk_samples), cnt * 8, cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        e->d2h += cnt * 8;
        return PAPR_OK;
    };
    rc = host_pass1_driver(e, g, budget / 8 / PAPR_SEQ_TILE + 2, true, exact ? SEQ_FULL : SEQ_NONE, &exact_done,
                           [&](const ChunkWork &work) { return stream_fd_chunks(e, fd, &g, &sc, budget, work); }, fetch);
    if (rc) return rc;
    out->sum_path = exact_done ? 3u : 0u;
    papr_launch_levels(&e->d_out->local, 1, e->tables(graph), graph, &e->d_out->merged, &e->d_out->lv,
                       &e->d_out->counts[PAPR_MAX_LEVELS], e->stream);
    e->launches += 1;
    auto ccdf_pass = [&]() -> int { // the resident chunks, one launch each
        int r;
        if ((r = hist_begin(e))) return r;
        for (size_t c = 0; c < sc.dev.size(); ++c) {
            const u64 off = c * sc.chunk_samples, m = std::min<u64>(sc.chunk_samples, g.n - off);
            if (m && (r = hist_chunk(e, sc.dev[c], m, false))) return r;
        }
        return hist_end(e);
    };
    if ((rc = ccdf_pass())) return rc;
    if ((rc = enqueue_fetch(e))) return rc;
    CU(cudaEventRecord(e->ev_end, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    stats_to_host(e->h_out->o.merged, &out->stats);
    if (std::isnan(out->stats.sum) && g.n) { // sign of the reference's printed "nan": the first NaN that entered the sum
        CU(cudaMemsetAsync(e->d_nan_idx, 0xff, sizeof(u64), e->stream));
        for (size_t c = 0; c < sc.dev.size(); ++c) {
            const u64 off = c * sc.chunk_samples, m = std::min<u64>(sc.chunk_samples, g.n - off);
            if (m) papr_launch_find_nan(sc.dev[c], m, off, e->d_nan_idx, e->num_sms * 8, e->stream);
        }
        CU(cudaMemcpyAsync(&e->h_out->nan_idx, e->d_nan_idx, sizeof(u64), cudaMemcpyDeviceToHost, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        const u64 k = e->h_out->nan_idx;
        if (k != ~0ull) {
            float pair[2];
            if ((rc = fetch(k, 1, pair))) return rc;
            const bool neg = std::isnan(pair[1]) ? std::signbit(pair[1]) : std::signbit(pair[0]);
            out->stats.sum = std::copysign(std::fabs(out->stats.sum), neg ? -1.0 : 1.0);
        }
    }
    if (collect(e, graph, out, true)) { // device level table != host libm's: redo with the host's
        if ((rc = upload_levels(e, out->level, out->nlevels, out->stats.peak))) return rc;
        if ((rc = enqueue_reset(e))) return rc;
        if ((rc = ccdf_pass())) return rc;
        if ((rc = enqueue_fetch(e))) return rc;
        CU(cudaEventRecord(e->ev_end, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        if (e->h_out->o.counts[PAPR_MAX_LEVELS] != 0) return fail(e, PAPR_ERR_INTERNAL, "exact CCDF pass reported a miss");
        for (int j = 0; j < out->nlevels; ++j) out->level_count[j] = (int64_t)e->h_out->o.counts[j];
    }
    finish_timing(e, out);
    return PAPR_OK;
}

static int reopen_direct(int fd)
{
    char p[64];
    snprintf(p, sizeof(p), "/proc/self/fd/%d", fd);
    return ::open(p, O_RDONLY | O_DIRECT);
}

// The capture behind a file descriptor: a regular file is read with pread() chunk by chunk (page cache ->
// pinned staging, no mapping); anything else is streamed once (analyze_stream).
extern "C" int papr_analyze_fd(papr_engine *e, int fd, int graph, papr_result *out)
{
    if (!e || fd < 0 || !out) return PAPR_ERR_ARG;
    cudaSetDevice(e->device);
    struct stat sb;
    if (fstat(fd, &sb) != 0) return fail(e, PAPR_ERR_IO, "fstat failed");
    if (!S_ISREG(sb.st_mode)) return analyze_stream(e, fd, graph, out);
    HostSource src;
    src.fd = fd;
    src.bytes = (u64)sb.st_size;
    posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
    src.dfd = e->o_direct ? reopen_direct(fd) : -1; // (tmpfs and some overlays refuse O_DIRECT: buffered reads then)
    const int rc = analyze_source(e, src, graph, out);
    if (src.dfd >= 0) ::close(src.dfd);
    return rc;
}

extern "C" int papr_analyze_file(papr_engine *e, const char *path, int graph, papr_result *out)
{
    if (!e || !path || !out) return PAPR_ERR_ARG;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(e, PAPR_ERR_IO, std::string("cannot open ") + path);
    const int rc = papr_analyze_fd(e, fd, graph, out);
    ::close(fd);
    return rc;
}

// ------------------------------------------------------------------------------------------------
// one capture over several GPUs of this box: byte-range shards, one engine + one host thread per GPU,
// pass-1 states folded in range order on the host (deterministic, first occurrence kept), the exact
// sequential sum chained across the shards, ONE all-reduce (NCCL over NVLink) of the level counts.
// ------------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <nccl.h> // types only: the library is dlopen()ed so that single-GPU users do not need it

namespace {

// Bind the calling thread to the CPUs next to a GPU (sysfs local_cpulist of its PCI function), so that
// the pinned staging slots it allocates and the helper threads it starts live on the GPU's NUMA
// node.  Best effort: returns false and changes nothing when the topology cannot be read.
static bool bind_thread_near_device(int device, cpu_set_t *saved)
{
    char busid[32] = {0};
    if (cudaDeviceGetPCIBusId(busid, (int)sizeof(busid), device) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char *p = busid; *p; ++p) *p = (char)tolower(*p);
    const std::string path = std::string("/sys/bus/pci/devices/") + busid + "/local_cpulist";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {0};
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return false;
    cpu_set_t want;
    CPU_ZERO(&want);
    int ncpu = 0;
    for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k < 1) continue;
        if (k == 1) b = a;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &want); ++ncpu; }
    }
    if (ncpu == 0) return false;
    if (pthread_getaffinity_np(pthread_self(), sizeof(*saved), saved) != 0) return false;
    cpu_set_t both;
    CPU_AND(&both, &want, saved); // never widen what the user (taskset, cgroup) allowed
    if (CPU_COUNT(&both) == 0 || CPU_EQUAL(&both, saved)) return false;
    return pthread_setaffinity_np(pthread_self(), sizeof(both), &both) == 0;
}

class Barrier {
public:
    explicit Barrier(int n) : n_(n), count_(0), gen_(0) {}
    void wait()
    {
        std::unique_lock<std::mutex> l(m_);
        const u64 g = gen_;
        if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
        else cv_.wait(l, [&] { return gen_ != g; });
    }
private:
    std::mutex m_;
    std::condition_variable cv_;
    int n_, count_;
    u64 gen_;
};

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && AllReduce && CommDestroy;
    }
};

} // namespace

struct papr_multi {
    int n = 0;
    std::vector<int> dev;
    std::vector<papr_engine *> eng;
    NcclApi nccl;
    std::vector<ncclComm_t> comms;
    std::thread nccl_init;
    bool nccl_ok = false;
    std::string nccl_note, err;
};

extern "C" const char *papr_multi_last_error(const papr_multi *m) { return m ? m->err.c_str() : g_create_error.c_str(); }

extern "C" const char *papr_multi_exchange(const papr_multi *m)
{
    return !m ? "" : (m->nccl_ok ? "nccl" : m->nccl_note.c_str());
}

extern "C" void papr_multi_destroy(papr_multi *m)
{
    if (!m) return;
    if (m->nccl_init.joinable()) m->nccl_init.join();
    if (m->nccl_ok) for (auto c : m->comms) m->nccl.CommDestroy(c);
    for (auto e : m->eng) papr_engine_destroy(e);
    delete m;
}

extern "C" int papr_multi_create(int ndev, const int *devices, papr_multi **out)
{
    if (!out || ndev < 1 || ndev > 64) return PAPR_ERR_ARG;
    *out = nullptr;
    papr_multi *m = new papr_multi();
    m->n = ndev;
    bool distinct = true;
    for (int r = 0; r < ndev; ++r) {
        m->dev.push_back(devices ? devices[r] : r);
        for (int q = 0; q < r; ++q) distinct &= m->dev[q] != m->dev[r];
    }
    // The one exchange of a single-process run is 16 KB of counts that every shard has just copied to
    // the host anyway, so the default folds them there (deterministic, ~1 us).  PAPR_B200_EXCHANGE=nccl
    // does it with one ncclAllReduce over NVLink instead; measured here ncclCommInitAll alone costs
    // 19-54 s (system NCCL 2.27.3, 2 GPUs) against 0.09 s for the whole analysis, hence opt-in.  The
    // communicator is set up beside engine creation and the first transfers.
    const char *xch = getenv("PAPR_B200_EXCHANGE");
    if (!(xch && std::string(xch) == "nccl")) {
        m->nccl_note = ndev > 1 ? "host" : "none (single shard)";
    } else if (ndev > 1 && distinct) {
        m->nccl_init = std::thread([m] {
            if (!m->nccl.load()) { m->nccl_note = "host (libnccl.so.2 not found)"; return; }
            m->comms.resize(m->n);
            ncclResult_t r = m->nccl.CommInitAll(m->comms.data(), m->n, m->dev.data());
            if (r != ncclSuccess) {
                m->nccl_note = std::string("host (ncclCommInitAll: ") + (m->nccl.GetErrorString ? m->nccl.GetErrorString(r) : "?") + ")";
                m->comms.clear();
                return;
            }
            m->nccl_ok = true;
        });
    } else {
        m->nccl_note = ndev > 1 ? "host (virtual shards on one device)" : "none (single shard)";
    }
    for (int r = 0; r < ndev; ++r) {
        papr_engine *e = nullptr;
        int rc = papr_engine_create(m->dev[r], &e);
        if (rc != PAPR_OK) {
            papr_multi_destroy(m);
            return rc;
        }
        e->host_share = ndev; // the helper threads of all shards share the host's cores and memory bandwidth
        m->eng.push_back(e);
    }
    *out = m;
    return PAPR_OK;
}

extern "C" int papr_multi_set(papr_multi *m, const char *name, double value)
{
    if (!m) return PAPR_ERR_ARG;
    for (auto e : m->eng) {
        int rc = papr_engine_set(e, name, value);
        if (rc) { m->err = e->err; return rc; }
    }
    return PAPR_OK;
}

static int multi_analyze_source(papr_multi *m, const HostSource &whole, int graph, papr_result *out)
{
    graph = graph ? 1 : 0;
    const int N = m->n;
    const u64 bytes = whole.bytes, nfloats = bytes / 4, npairs = nfloats / 2;
    const bool tail = (nfloats & 1) != 0;
    float tail_pair[2] = {0.f, 0.f};
    if (tail) {
        if (!whole.read(&tail_pair[0], 4 * (nfloats - 1), 4)) { m->err = "read failed"; return PAPR_ERR_IO; }
        tail_pair[1] = stale_q_of(whole); // needs the WHOLE capture: the stale Q may sit in another shard
    }
    // shard = whole chunks (so chunk, batch and tile boundaries coincide with shard boundaries)
    const u64 chunk_samples = m->eng[0]->chunk_bytes / 8;
    u64 per = (npairs + N - 1) / N;
    per = std::max<u64>(chunk_samples, (per + chunk_samples - 1) / chunk_samples * chunk_samples);
    const int tail_rank = (int)std::min<u64>((u64)N - 1, npairs / per);

    reset_result(out);
    out->mode_used = PAPR_MODE_TWO_PASS;
    std::vector<papr_stats> st(N);
    std::vector<u64> ns(N, 0);
    std::vector<int> rcs(N, PAPR_OK);
    std::atomic<bool> failed(false);
    bool exact = m->eng[0]->exact_sum != 0, exact_ok = false, use_nccl = false;
    papr_stats merged;
    int L = 0;
    Barrier bar(N);
    auto t_begin = std::chrono::steady_clock::now();
    const bool trace = getenv("PAPR_B200_TRACE") != nullptr;
    auto mark = [&](int r, const char *what) {
        if (trace && r == 0)
            fprintf(stderr, "papr_b200 trace: %-28s %9.3f ms\n", what,
                    std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    };

    auto rank_main = [&](int r) {
        papr_engine *e = m->eng[r];
        const u64 lo = std::min(npairs, (u64)r * per), hi = std::min(npairs, (u64)(r + 1) * per);
        auto step = [&](int rc) { if (rc < 0 && !failed.exchange(true)) { rcs[r] = rc; } return rc; };
        begin_analysis(e);
        cpu_set_t saved_cpus;
        const bool bound = N > 1 && bind_thread_near_device(e->device, &saved_cpus);
        // pass 1 while the shard streams in
        cudaSetDevice(e->device);
        if (!failed) step(host_stream_stats(e, whole.slice(lo * 8, (hi - lo) * 8), lo, &ns[r], (tail && r == tail_rank) ? tail_pair : nullptr));
        if (!failed) {
            cudaMemcpyAsync(&e->h_out->o.local, &e->d_out->local, sizeof(PaprDevStats), cudaMemcpyDeviceToHost, e->stream);
            if (cudaStreamSynchronize(e->stream) != cudaSuccess) step(fail(e, PAPR_ERR_CUDA, cudaGetErrorString(cudaGetLastError())));
            stats_to_host(e->h_out->o.local, &st[r]);
            if (!failed) step(fix_nan_sign(e, e->d_buf, ns[r], lo, &st[r]));
        }
        if (!failed && exact) step(seq_tile_sums(e, e->d_buf, ns[r]));
        mark(r, "pass 1 + tile sums (rank 0)");
        bar.wait(); // ---- A: every shard's pass-1 state and tile sums are on the host
        mark(r, "barrier A");
        if (r == 0 && !failed) {
            merged = st[0];
            for (int q = 1; q < N; ++q) papr_stats_merge(&merged, &st[q]);
            if (exact && std::isfinite(merged.sum)) {
                double pre = 0.0;
                exact_ok = true;
                for (int q = 0; q < N && exact_ok; ++q) {
                    const u64 nt = (ns[q] + PAPR_SEQ_TILE - 1) / PAPR_SEQ_TILE;
                    if (seq_classify(m->eng[q]->h_tile_sum, nt, &pre, m->eng[q]->h_tile_code)) exact_ok = false;
                }
            }
        }
        bar.wait(); // ---- B
        if (!failed && exact_ok) step(seq_tile_runs(e, e->d_buf, ns[r]));
        bar.wait(); // ---- C: every shard's tile runs are on the host
        mark(r, "barrier C (tile runs)");
        if (r == 0 && !failed) {
            if (exact_ok) {
                double s = 0.0;
                for (int q = 0; q < N; ++q)
                    if (seq_chain_resident(m->eng[q], m->eng[q]->d_buf, ns[q], &s) < 0) { failed = true; rcs[0] = PAPR_ERR_CUDA; }
                merged.sum = s;
            }
            out->stats = merged;
            L = papr_result_finish(out, graph); // host libm: avg, papr, levels
            mark(r, "chained sum + host levels");
            if (m->nccl_init.joinable()) m->nccl_init.join();
            use_nccl = m->nccl_ok;
            mark(r, "nccl init joined");
        }
        bar.wait(); // ---- D: levels known
        cudaSetDevice(e->device);
        if (!failed && L > 0) {
            int rc = upload_levels(e, out->level, L, merged.peak);
            if (!rc) rc = enqueue_reset(e);
            if (!rc) rc = enqueue_hist_exact(e, e->d_buf, ns[r], true);
            step(rc);
        }
        bool nccl_called = false;
        if (L > 0 && use_nccl) { // every rank must take part, even one that failed locally
            ncclResult_t nr = m->nccl.AllReduce(e->d_out->counts, e->d_out->counts, PAPR_MAX_LEVELS + 1, ncclUint64, ncclSum,
                                                m->comms[r], e->stream);
            nccl_called = nr == ncclSuccess;
            if (!nccl_called) step(fail(e, PAPR_ERR_CUDA, "ncclAllReduce failed"));
        }
        if (L > 0) {
            enqueue_fetch(e);
            cudaEventRecord(e->ev_end, e->stream);
            if (cudaStreamSynchronize(e->stream) != cudaSuccess) step(fail(e, PAPR_ERR_CUDA, "synchronise failed after the CCDF pass"));
        }
        mark(r, "CCDF pass + exchange (rank 0)");
        bar.wait(); // ---- E: counts on the host
        mark(r, "barrier E");
        if (bound) pthread_setaffinity_np(pthread_self(), sizeof(saved_cpus), &saved_cpus);
    };
    std::vector<std::thread> th;
    for (int r = 1; r < N; ++r) th.emplace_back(rank_main, r);
    rank_main(0);
    for (auto &t : th) t.join();
    if (failed) {
        for (int r = 0; r < N; ++r)
            if (rcs[r] < 0) { m->err = "shard " + std::to_string(r) + ": " + m->eng[r]->err; return rcs[r]; }
        m->err = "shard failure";
        return PAPR_ERR_CUDA;
    }
    u64 miss = 0;
    for (int j = 0; j < L; ++j) out->level_count[j] = 0;
    for (int q = 0; q < (use_nccl ? 1 : N) && L > 0; ++q) {
        const DevOut &o = m->eng[q]->h_out->o;
        for (int j = 0; j < L; ++j) out->level_count[j] += (int64_t)o.counts[j];
        miss += o.counts[PAPR_MAX_LEVELS];
    }
    if (miss) { m->err = "exact CCDF pass reported a miss"; return PAPR_ERR_INTERNAL; }
    for (int q = 0; q < N; ++q) {
        out->kernel_launches += m->eng[q]->launches;
        out->h2d_bytes += m->eng[q]->h2d;
        out->d2h_bytes += m->eng[q]->d2h;
    }
    out->device_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return PAPR_OK;
}

extern "C" int papr_multi_analyze_host(papr_multi *m, const void *image, uint64_t bytes, int graph, papr_result *out)
{
    if (!m || !out || (bytes && !image)) return PAPR_ERR_ARG;
    cudaSetDevice(m->eng[0]->device);
    return multi_analyze_source(m, memory_source(image, bytes), graph, out);
}

// a regular file is sharded by byte range over the GPUs; a pipe is read once by the first engine (the pipe,
// not the GPU, is the limit there)
extern "C" int papr_multi_analyze_fd(papr_multi *m, int fd, int graph, papr_result *out)
{
    if (!m || fd < 0 || !out) return PAPR_ERR_ARG;
    struct stat sb;
    if (fstat(fd, &sb) != 0) { m->err = "fstat failed"; return PAPR_ERR_IO; }
    if (!S_ISREG(sb.st_mode)) {
        const int rc = papr_analyze_fd(m->eng[0], fd, graph, out);
        if (rc) m->err = m->eng[0]->err;
        return rc;
    }
    HostSource src;
    src.fd = fd;
    src.bytes = (u64)sb.st_size;
    posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
    src.dfd = m->eng[0]->o_direct ? reopen_direct(fd) : -1;
    const int rc = multi_analyze_source(m, src, graph, out);
    if (src.dfd >= 0) ::close(src.dfd);
    return rc;
}

extern "C" int papr_multi_analyze_file(papr_multi *m, const char *path, int graph, papr_result *out)
{
    if (!m || !path || !out) return PAPR_ERR_ARG;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) { m->err = std::string("cannot open ") + path; return PAPR_ERR_IO; }
    const int rc = papr_multi_analyze_fd(m, fd, graph, out);
    ::close(fd);
    return rc;
}
