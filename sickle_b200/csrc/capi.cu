// capi.cu -- host runtime behind include/sickle_b200.h: slots, pinned double buffering, streams,
// kernel launches.  The only public symbols are the extern "C" entry points of the header.
//
// Pipeline per slot (one CUDA stream per slot; slots overlap each other's copies and kernels):
//   H2D(in) -> kernels -> D2H(summary);  sk_wait: wait summary, D2H(out streams, exact sizes).
// Kernels of a batch are either
//   fused   : kf_fused (parse+trim+route+emit in one pass) + kf_finalize      [short records, input order]
//   two files: kf_fused<CH,1> (verdicts) -> kf2_between -> kf_fused<CH,2> (route + emit) -> kf2_finalize
//   ordered : kf_fused<CH,3> (index + verdicts) -> kfo_offsets -> kf_fused<CH,4> (emit queue by queue) -> kf_finalize
//                                                                               [-a N, single end, N <= 32]
//   hybrid  : kf_fused<CH,3> (line index + verdicts, one or two inputs) -> k2_trim_route<true> -> K3 -> k_finalize   [-a N otherwise]
//   general : K1 line index (per input) -> K2 trim+route+scan (one or two kernels) -> K3 emit -> k_finalize   [everything]
// The fused kernel flags what it cannot do exactly (long records, data errors, ...); the batch is then
// re-run on the general path before the result is returned, and the context backs off for a while.
// This replaces the reference's two-stage overlap (detached output thread while the main thread
// reads the next batch: src/trim_single.cpp:336-339, src/trim_paired.cpp:444-458).
#include "../../include/sickle_b200.h"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "k1_index.cuh"
#include "k2_trim.cuh"
#include "k3_emit.cuh"
#include "kf_fused.cuh"
#include "sk_device.cuh"

namespace {

thread_local char g_err[512] = "";

void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

#define SK_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            set_err("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                    cudaGetErrorString(e__));                                                \
            return SK_E_CUDA;                                                                \
        }                                                                                    \
    } while (0)

// SICKLE_B200_DEBUG_SYNC=1: synchronise after every kernel and name the one that faulted
#define SK_DEBUG_SYNC(st, name)                                                                     \
    do {                                                                                            \
        static const bool dbg__ = getenv("SICKLE_B200_DEBUG_SYNC") != nullptr;                      \
        if (dbg__) {                                                                                \
            cudaError_t e__ = cudaStreamSynchronize(st);                                            \
            if (e__ == cudaSuccess) e__ = cudaGetLastError();                                       \
            if (e__ != cudaSuccess) {                                                               \
                set_err("kernel %s failed: %s", name, cudaGetErrorString(e__));                     \
                fprintf(stderr, "[sickle_b200] kernel %s failed: %s\n", name, cudaGetErrorString(e__)); \
                return SK_E_CUDA;                                                                   \
            }                                                                                       \
        }                                                                                           \
    } while (0)

constexpr uint64_t kPad = 64;           // readable padding after every device input buffer
constexpr uint64_t kMaxSlotBytes = (1ull << 31) - 4096;
constexpr int kFusedMinTile = sk::FusedCfg<3>::kTile;   // smallest tile of the instantiated configs

// what a batch was launched with (kept so that a fused batch can be re-run on the general path)
struct BatchArgs {
    const uint8_t *in[2] = {nullptr, nullptr};
    uint32_t first[2] = {0, 0};
    uint64_t n[2] = {0, 0};
    uint8_t *out[3] = {nullptr, nullptr, nullptr};
    uint64_t cap[3] = {0, 0, 0};
    cudaStream_t st = nullptr;
};

struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_done = nullptr;
    cudaEvent_t ev_stage[3] = {nullptr, nullptr, nullptr};   // after K1, after K2, after K3
    char *h_in[2] = {nullptr, nullptr};
    char *h_out[3] = {nullptr, nullptr, nullptr};
    uint8_t *d_in[2] = {nullptr, nullptr};
    uint8_t *d_out[3] = {nullptr, nullptr, nullptr};
    uint64_t out_cap[3] = {0, 0, 0};
    uint32_t *d_line_end[2] = {nullptr, nullptr};
    sk::RecDesc *d_desc[2] = {nullptr, nullptr};
    unsigned long long *d_status_k1[2] = {nullptr, nullptr};
    unsigned long long *d_status_k2 = nullptr;
    unsigned long long *d_status_f = nullptr;   // fused: [3][fused_tiles_cap] (newlines, main, singles); two files: [2 + 4]
    unsigned long long *d_verdict[2] = {nullptr, nullptr};   // fused, two files: one 8-byte entry per record and file
    uint8_t *d_nlsave[2] = {nullptr, nullptr};               // fused, two files / -a N: the first pass's newline positions, kFNlSlot bytes per tile
    uint32_t *d_tq = nullptr;                                // -a N: kept bytes per tile and queue, [tiles + 1][32]
    sk::Control *d_ctl = nullptr;
    sk::DevResult *d_res = nullptr;
    sk::DevResult *h_res = nullptr;
    uint32_t epoch = 0;
    bool busy = false;
    bool device_mode = false;
    bool last_fused = false;
    uint64_t up_lo[2] = {0, 0}, up_hi[2] = {0, 0};   // byte range already uploaded by sk_upload
    uint64_t base[2] = {0, 0};                       // buffer offset of DevInput.data for the batch
    uint32_t first[2] = {0, 0};                      // start - base (0..15)
    BatchArgs last;
    uint32_t launches = 0;
};

}  // namespace

struct sk_ctx {
    int device = 0;
    int sm_count = 148;
    uint64_t slot_bytes = 0;
    uint32_t line_cap = 0;       // lines per input
    uint32_t k1_tiles_cap = 0;
    uint32_t k2_tiles_cap = 0;
    uint32_t fused_tiles_cap = 0;
    uint32_t verdict_cap = 0;    // two files: records per file the verdict tables hold
    int n_inputs = 1;
    sk_params params{};
    sk::DevParams dev{};
    std::vector<Slot> slots;
    bool host_buffers = false;
    // path selection
    bool fused_eligible = false;   // mode / order the fused kernel supports
    bool hybrid_eligible = false;  // -a N on one input: the fused kernel's index + verdict pass, then k2_trim_route<true> and K3
    bool ordered_eligible = false; // -a N, N <= 32, single end: index + verdict pass, then the fused kernel's ordered emit (no K3)
    int fused_grid_ordered[5] = {0, 0, 0, 0, 0};
    int fused_ch = 7;              // 16-byte chunks per thread (3, 5, 7, 9 or 11): the tile size in use
    bool fused_ch_fixed = false;   // SICKLE_B200_FUSED_CH given: no adaptation
    int fused_ch_max = 9;          // lowered after a failed fused batch, raised again after a streak of good ones
    int fused_ok_streak = 0;
    int fused_grid_ch[5] = {0, 0, 0, 0, 0};   // persistent grid per tile size, index (CH - 3) / 2
    int fused_grid_pass1[5] = {0, 0, 0, 0, 0};   // two files, PASS 1 (no staging buffer: more CTAs per SM)
    int fused_backoff = 0;         // batches left on the general path after a fused failure
    int fused_fail_streak = 0;     // fused failures without a fused success in between: each doubles the back-off
    int k2_split = -1;             // SICKLE_B200_K2_SPLIT: 0 never, 1 always, unset = for long records and -a N
    bool long_records = false;     // the last general-path batch averaged 1.5 KB or more per record: K2 runs as two kernels
    uint64_t n_fused = 0, n_general = 0, n_rerun = 0;
};

namespace {

int make_dev_params(const sk_params &p, sk::DevParams &d) {
    // quality_constants {offset, min, max}: reference src/sickle.h:85-91
    switch (p.qualtype) {
        case SK_QUAL_SANGER: d.qoff = 33; d.qmin = 33; d.qmax = 126; break;
        case SK_QUAL_SOLEXA: d.qoff = 64; d.qmin = 58; d.qmax = 112; break;
        case SK_QUAL_ILLUMINA: d.qoff = 64; d.qmin = 64; d.qmax = 110; break;
        default: set_err("invalid qualtype %d", p.qualtype); return SK_E_ARG;
    }
    if (p.qual_threshold < 0 || p.length_threshold < 0) { set_err("thresholds must be >= 0"); return SK_E_ARG; }
    if (p.mode < SK_MODE_SE || p.mode > SK_MODE_PE_INTER_M) { set_err("invalid mode %d", p.mode); return SK_E_ARG; }
    d.qthr = p.qual_threshold;
    d.lthr = p.length_threshold;
    d.no_fiveprime = p.no_fiveprime ? 1 : 0;
    d.trunc_n = p.trunc_n ? 1 : 0;
    d.mode = p.mode;
    d.emu_threads = p.emulate_threads > 1 ? p.emulate_threads : 1;
    d.has_singles = p.has_singles ? 1 : 0;
    return SK_OK;
}

void free_slot(Slot &s) {
    for (int i = 0; i < 2; ++i) {
        if (s.h_in[i]) cudaFreeHost(s.h_in[i]);
        if (s.d_in[i]) cudaFree(s.d_in[i]);
        if (s.d_line_end[i]) cudaFree(s.d_line_end[i]);
        if (s.d_desc[i]) cudaFree(s.d_desc[i]);
        if (s.d_status_k1[i]) cudaFree(s.d_status_k1[i]);
    }
    for (int i = 0; i < 3; ++i) {
        if (s.h_out[i]) cudaFreeHost(s.h_out[i]);
        if (s.d_out[i]) cudaFree(s.d_out[i]);
    }
    if (s.d_status_k2) cudaFree(s.d_status_k2);
    if (s.d_status_f) cudaFree(s.d_status_f);
    for (auto &v : s.d_verdict) if (v) cudaFree(v);
    for (auto &v : s.d_nlsave) if (v) cudaFree(v);
    if (s.d_tq) cudaFree(s.d_tq);
    if (s.d_ctl) cudaFree(s.d_ctl);
    if (s.d_res) cudaFree(s.d_res);
    if (s.h_res) cudaFreeHost(s.h_res);
    if (s.ev_begin) cudaEventDestroy(s.ev_begin);
    if (s.ev_end) cudaEventDestroy(s.ev_end);
    if (s.ev_done) cudaEventDestroy(s.ev_done);
    for (auto &e : s.ev_stage) if (e) cudaEventDestroy(e);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Slot{};
}

int alloc_slot(sk_ctx *c, Slot &s, bool host_buffers) {
    const uint64_t sb = c->slot_bytes;
    SK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    SK_CUDA(cudaEventCreate(&s.ev_begin));
    SK_CUDA(cudaEventCreate(&s.ev_end));
    for (auto &e : s.ev_stage) SK_CUDA(cudaEventCreate(&e));
    SK_CUDA(cudaEventCreateWithFlags(&s.ev_done, cudaEventDisableTiming));
    for (int i = 0; i < c->n_inputs; ++i) {
        if (host_buffers) {
            // SICKLE_B200_WC_INPUT=1: write-combined pinned input slots (the host only ever writes them; the
            // copy engine's reads then skip the CPU caches).  Off by default: the command line's carry of an
            // unconsumed tail reads the slot back, which is slow on write-combined memory.
            static const bool wc = getenv("SICKLE_B200_WC_INPUT") && atoi(getenv("SICKLE_B200_WC_INPUT")) != 0;
            SK_CUDA(cudaHostAlloc((void **)&s.h_in[i], sb + kPad, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
            SK_CUDA(cudaMalloc((void **)&s.d_in[i], sb + kPad));
            SK_CUDA(cudaMemset(s.d_in[i], 0, sb + kPad));
        }
        SK_CUDA(cudaMalloc((void **)&s.d_line_end[i], (size_t)c->line_cap * sizeof(uint32_t) + 64));
        SK_CUDA(cudaMalloc((void **)&s.d_desc[i], ((size_t)c->line_cap / 4 + 1) * sizeof(sk::RecDesc)));
        SK_CUDA(cudaMalloc((void **)&s.d_status_k1[i], (size_t)c->k1_tiles_cap * 8 * sk::kWideStatusStride));
        SK_CUDA(cudaMemset(s.d_status_k1[i], 0, (size_t)c->k1_tiles_cap * 8 * sk::kWideStatusStride));
    }
    SK_CUDA(cudaMalloc((void **)&s.d_status_k2, (size_t)c->k2_tiles_cap * 8 * sk::kMaxStreams));
    SK_CUDA(cudaMemset(s.d_status_k2, 0, (size_t)c->k2_tiles_cap * 8 * sk::kMaxStreams));
    if (c->fused_eligible || c->hybrid_eligible) {   // (-a N on two files: the index pass walks a newline chain per file)
        const size_t chains = c->n_inputs == 2 ? 6 : 3;   // two files: a newline chain and two output chains per file
        SK_CUDA(cudaMalloc((void **)&s.d_status_f, (size_t)c->fused_tiles_cap * 8 * chains * sk::kWideStatusStride));
        SK_CUDA(cudaMemset(s.d_status_f, 0, (size_t)c->fused_tiles_cap * 8 * chains * sk::kWideStatusStride));
        if (c->n_inputs == 2 && c->fused_eligible) {
            for (auto &v : s.d_verdict) SK_CUDA(cudaMalloc((void **)&v, (size_t)c->verdict_cap * 8));
            for (auto &v : s.d_nlsave) SK_CUDA(cudaMalloc((void **)&v, (size_t)c->fused_tiles_cap * sk::kFNlSlot));
        }
        if (c->ordered_eligible) {
            SK_CUDA(cudaMalloc((void **)&s.d_nlsave[0], (size_t)c->fused_tiles_cap * sk::kFNlSlot));
            SK_CUDA(cudaMalloc((void **)&s.d_tq, ((size_t)c->fused_tiles_cap + 1 + c->fused_tiles_cap / sk::kFTqGroup + 2) * 32 * sizeof(uint32_t)));
        }
    }
    if (host_buffers) {
        // stream capacities: an output stream never exceeds the bytes of the inputs feeding it
        const int mode = c->params.mode;
        uint64_t cap[3] = {sb, 0, 0};
        if (mode == SK_MODE_PE_2FILE) { cap[1] = sb; cap[2] = c->params.has_singles ? 2 * sb : 0; }
        if (mode == SK_MODE_PE_INTER) cap[2] = c->params.has_singles ? sb : 0;
        for (int k = 0; k < 3; ++k) {
            s.out_cap[k] = cap[k];
            if (!cap[k]) continue;
            SK_CUDA(cudaHostAlloc((void **)&s.h_out[k], cap[k] + kPad, cudaHostAllocDefault));
            SK_CUDA(cudaMalloc((void **)&s.d_out[k], cap[k] + kPad));
        }
    }
    SK_CUDA(cudaMalloc((void **)&s.d_ctl, sizeof(sk::Control)));
    sk::Control z;
    memset(&z, 0, sizeof z);
    z.err_key = sk::kNoError;
    SK_CUDA(cudaMemcpy(s.d_ctl, &z, sizeof z, cudaMemcpyHostToDevice));
    SK_CUDA(cudaMalloc((void **)&s.d_res, sizeof(sk::DevResult)));
    SK_CUDA(cudaHostAlloc((void **)&s.h_res, sizeof(sk::DevResult), cudaHostAllocDefault));
    return SK_OK;
}

void make_inputs(const sk_ctx *c, const Slot &s, const BatchArgs &a, sk::DevInput di[2], sk::OutPtrs &op) {
    for (int i = 0; i < 2; ++i) {
        di[i].data = a.in[i];
        di[i].first = a.first[i];
        di[i].nbytes = (uint32_t)a.n[i];
        di[i].line_end = s.d_line_end[i];
        di[i].line_cap = c->line_cap;
    }
    if (c->n_inputs == 1) { di[1].data = nullptr; di[1].first = 0; di[1].nbytes = 0; di[1].line_end = nullptr; di[1].line_cap = 0; }
    for (int k = 0; k < 3; ++k) { op.p[k] = a.out[k]; op.cap[k] = a.out[k] ? a.cap[k] : 0; }
}

int next_epoch(sk_ctx *c, Slot &s, cudaStream_t st) {
    s.epoch += 1;
    if ((s.epoch & (uint32_t)sk::kEpochMask) == 0) {  // epoch tag wrapped: clear the status words once
        for (int i = 0; i < c->n_inputs; ++i) SK_CUDA(cudaMemsetAsync(s.d_status_k1[i], 0, (size_t)c->k1_tiles_cap * 8 * sk::kWideStatusStride, st));
        SK_CUDA(cudaMemsetAsync(s.d_status_k2, 0, (size_t)c->k2_tiles_cap * 8 * sk::kMaxStreams, st));
        if (s.d_status_f) SK_CUDA(cudaMemsetAsync(s.d_status_f, 0, (size_t)c->fused_tiles_cap * 8 * (c->n_inputs == 2 ? 6 : 3) * sk::kWideStatusStride, st));
        s.epoch += 1;
    }
    return SK_OK;
}

// General path: K1 (per input) -> K2 -> K3 -> finalize.
int launch_general(sk_ctx *c, Slot &s, const BatchArgs &a) {
    cudaStream_t st = a.st;
    if (int rc = next_epoch(c, s, st)) return rc;
    sk::DevInput di[2];
    sk::OutPtrs op;
    make_inputs(c, s, a, di, op);
    s.launches = 0;
    SK_CUDA(cudaEventRecord(s.ev_begin, st));
    const int resident = c->sm_count * 8;
    for (int i = 0; i < c->n_inputs; ++i) {
        const uint32_t tiles = (uint32_t)((a.n[i] + sk::kK1TileBytes - 1) / sk::kK1TileBytes);
        if (!tiles) continue;
        const int grid = tiles < (uint32_t)resident ? (int)tiles : resident;
        sk::k1_line_index<<<grid, sk::kK1Threads, 0, st>>>(di[i], s.d_ctl, i, s.d_status_k1[i], tiles, s.epoch);
        s.launches++;
        SK_DEBUG_SYNC(st, "k1_line_index");
    }
    SK_CUDA(cudaEventRecord(s.ev_stage[0], st));
    // units <= bytes / 4 (a line is at least its '\n'); usually ~bytes/325
    const uint64_t max_units = (a.n[0] + a.n[1]) / 4 + 1;
    const uint64_t tiles = (max_units + sk::kK2UnitsPerTile - 1) / sk::kK2UnitsPerTile;
    const int grid = tiles < (uint64_t)resident ? (int)tiles : resident;
    // K2 as two kernels: long records (no tile waits behind a 20 kb read), and -a N order, where the lanes of the
    // one-kernel form trim records N apart (same-GPU A/B, -a 8, 1 M reads: 0.654 -> 0.632 ms; input order: +1 %)
    const bool split = c->k2_split < 0 ? (c->long_records || c->dev.emu_threads > 1) : c->k2_split != 0;
    if (split) {   // trimming on its own (no tile waits for another), then routing + scan from the verdicts
        sk::k2_trim_only<<<resident, sk::kK2Threads, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1],
                                                              c->long_records ? sk::kK2aLongUnitsPerTicket : 32u);
        SK_DEBUG_SYNC(st, "k2_trim_only");
        sk::k2_trim_route<true><<<grid, sk::kK2Threads, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1],
                                                                 s.d_status_k2, c->k2_tiles_cap, s.epoch);
        s.launches++;
    } else {
        sk::k2_trim_route<false><<<grid, sk::kK2Threads, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1],
                                                                  s.d_status_k2, c->k2_tiles_cap, s.epoch);
    }
    SK_DEBUG_SYNC(st, "k2_trim_route");
    SK_CUDA(cudaEventRecord(s.ev_stage[1], st));
    if (c->long_records) sk::k3_emit_long<<<resident, 256, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1], op);
    else sk::k3_emit<<<resident, sk::kK3Threads, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1], op);
    SK_DEBUG_SYNC(st, "k3_emit");
    SK_CUDA(cudaEventRecord(s.ev_stage[2], st));
    sk::k_finalize<<<1, 32, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, op, s.d_res);
    s.launches += 3;
    SK_DEBUG_SYNC(st, "k_finalize");
    SK_CUDA(cudaEventRecord(s.ev_end, st));
    SK_CUDA(cudaGetLastError());
    s.last_fused = false;
    c->n_general++;
    return SK_OK;
}

template <int CH>
int launch_fused_ch(sk_ctx *c, Slot &s, const BatchArgs &a, const sk::DevInput &di, const sk::OutPtrs &op) {
    using Cfg = sk::FusedCfg<CH>;
    cudaStream_t st = a.st;
    const uint32_t tiles = (uint32_t)((a.n[0] + Cfg::kTile - 1) / Cfg::kTile);
    if (tiles) {
        const int full = c->fused_grid_ch[(CH - 3) / 2];
        const int grid = tiles < (uint32_t)full ? (int)tiles : full;
        sk::kf_fused<CH><<<grid, sk::kFThreads, Cfg::kSmem, st>>>(di, c->dev, s.d_ctl, op, s.d_status_f,
                                                                 s.d_status_f + (size_t)c->fused_tiles_cap * sk::kWideStatusStride,
                                                                 c->fused_tiles_cap * sk::kWideStatusStride,
                                                                 tiles, s.epoch);
        s.launches++;
        SK_DEBUG_SYNC(st, "kf_fused");
    }
    return SK_OK;
}

// Two files: both passes of kf_fused over the tiles of both inputs (see kf_fused.cuh), the small kernel between them.
template <int CH>
int launch_fused_two_ch(sk_ctx *c, Slot &s, const BatchArgs &a, const sk::DevInput di[2], const sk::OutPtrs &op) {
    using Cfg = sk::FusedCfg<CH>;
    cudaStream_t st = a.st;
    const uint32_t tiles_a = (uint32_t)((a.n[0] + Cfg::kTile - 1) / Cfg::kTile), tiles_b = (uint32_t)((a.n[1] + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t tiles = tiles_a + tiles_b;
    const uint32_t stride = c->fused_tiles_cap * sk::kWideStatusStride;
    if (tiles) {
        const int full = c->fused_grid_ch[(CH - 3) / 2];
        const int grid = tiles < (uint32_t)full ? (int)tiles : full;
        const int grid1 = tiles < (uint32_t)c->fused_grid_pass1[(CH - 3) / 2] ? (int)tiles : c->fused_grid_pass1[(CH - 3) / 2];
        sk::kf_fused<CH, 1><<<grid1, sk::kFThreads, Cfg::kSmemPass1, st>>>(di[0], c->dev, s.d_ctl, op, s.d_status_f, s.d_status_f + 2 * (size_t)stride,
                                                                    stride, tiles, s.epoch, di[1], tiles_b, s.d_verdict[0], s.d_verdict[1], c->verdict_cap,
                                                                    s.d_nlsave[0], s.d_nlsave[1]);
        SK_DEBUG_SYNC(st, "kf_fused pass 1");
        sk::kf2_between<<<1, 32, 0, st>>>(s.d_ctl);
        if (int rc = next_epoch(c, s, st)) return rc;   // the newline chains are walked again
        sk::kf_fused<CH, 2><<<grid, sk::kFThreads, Cfg::kSmemTwoFile, st>>>(di[0], c->dev, s.d_ctl, op, s.d_status_f, s.d_status_f + 2 * (size_t)stride,
                                                                           stride, tiles, s.epoch, di[1], tiles_b, s.d_verdict[0], s.d_verdict[1], c->verdict_cap,
                                                                    s.d_nlsave[0], s.d_nlsave[1]);
        SK_DEBUG_SYNC(st, "kf_fused pass 2");
        s.launches += 3;
    }
    return SK_OK;
}

template <int CH>
int setup_fused_ch(sk_ctx *c) {
    using Cfg = sk::FusedCfg<CH>;
    SK_CUDA(cudaFuncSetAttribute(sk::kf_fused<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    int per_sm = 0;
    SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sk::kf_fused<CH>, sk::kFThreads, Cfg::kSmem));
    if (per_sm < 1) { set_err("fused kernel does not fit on this device"); return SK_E_CUDA; }
    if constexpr (CH != 11) if (c->n_inputs == 2) {   // (two files: tiles of 18 / 25 / 32 KB)
        SK_CUDA(cudaFuncSetAttribute(sk::kf_fused<CH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemPass1));
        int p1 = 0;
        SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p1, sk::kf_fused<CH, 1>, sk::kFThreads, Cfg::kSmemPass1));
        if (p1 < 1) { set_err("fused kernel does not fit on this device"); return SK_E_CUDA; }
        c->fused_grid_pass1[(CH - 3) / 2] = p1 * c->sm_count;
        SK_CUDA(cudaFuncSetAttribute(sk::kf_fused<CH, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemTwoFile));
        int p2 = 0;
        SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, sk::kf_fused<CH, 2>, sk::kFThreads, Cfg::kSmemTwoFile));
        if (p2 < 1) { set_err("fused kernel does not fit on this device"); return SK_E_CUDA; }
        if (p2 < per_sm) per_sm = p2;   // one grid size for both passes
    }
    if constexpr (CH != 11) if (c->hybrid_eligible) {   // the index + verdict pass stages nothing, like PASS 1
        SK_CUDA(cudaFuncSetAttribute(sk::kf_fused<CH, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemPass1));
        int p3 = 0;
        SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p3, sk::kf_fused<CH, 3>, sk::kFThreads, Cfg::kSmemPass1));
        if (p3 < 1) { set_err("fused kernel does not fit on this device"); return SK_E_CUDA; }
        c->fused_grid_pass1[(CH - 3) / 2] = p3 * c->sm_count;
        if (c->ordered_eligible) {
            SK_CUDA(cudaFuncSetAttribute(sk::kf_fused<CH, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemOrdered));
            int p4 = 0;
            SK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p4, sk::kf_fused<CH, 4>, sk::kFThreads, Cfg::kSmemOrdered));
            if (p4 < 1) { set_err("fused kernel does not fit on this device"); return SK_E_CUDA; }
            c->fused_grid_ordered[(CH - 3) / 2] = p4 * c->sm_count;
        }
    }
    c->fused_grid_ch[(CH - 3) / 2] = per_sm * c->sm_count;
    return SK_OK;
}

int setup_fused(sk_ctx *c) {
    if (c->fused_ch != 3 && c->fused_ch != 5 && c->fused_ch != 7 && c->fused_ch != 9 && c->fused_ch != 11) c->fused_ch = 7;
    if (int rc = setup_fused_ch<3>(c)) return rc;
    if (int rc = setup_fused_ch<5>(c)) return rc;
    if (int rc = setup_fused_ch<7>(c)) return rc;
    if (int rc = setup_fused_ch<9>(c)) return rc;
    return setup_fused_ch<11>(c);
}

// Tile size follows the records.  A tile may hold at most 128 records (two staging lanes per record),
// and larger tiles amortise the per-tile work better (measured on 150-base reads, ms per 1M reads:
// CH 5: 0.410, 7: 0.360, 9: 0.340, 11: 0.354).  After every batch the average record size picks the
// largest tile that stays under ~112 records; records too short even for the smallest tile go
// straight to the general path.  A failed fused batch lowers the ceiling for a while.
uint32_t fused_tile_bytes(int ch) { return (uint32_t)sk::kFTileThreads * 16u * (uint32_t)ch; }

void adapt_fused(sk_ctx *c, const sk::DevResult &r, bool failed) {
    if (!(c->fused_eligible || c->hybrid_eligible) || c->fused_ch_fixed) return;
    if (failed) {
        c->fused_ok_streak = 0;
        if ((r.index_overflow & 8u) && c->fused_ch > 3) {   // too many records per tile: smaller tiles at once, no back-off
            c->fused_ch_max = c->fused_ch - 2;
            c->fused_ch = c->fused_ch_max;
            c->fused_backoff = 0;
        }
        return;
    }
    c->fused_fail_streak = 0;
    if (++c->fused_ok_streak >= 64 && c->fused_ch_max < 9) { c->fused_ch_max += 2; c->fused_ok_streak = 0; }
    const uint64_t records = r.records[0] + r.records[1];
    if (records < 64) return;
    const uint64_t avg = (r.consumed[0] + r.consumed[1]) / records;
    for (int ch = c->fused_ch_max; ch >= 3; ch -= 2)
        if ((uint64_t)fused_tile_bytes(ch) <= avg * 112u) { c->fused_ch = ch; return; }
    c->fused_ch = 3;
    if (c->fused_backoff < 4) c->fused_backoff = 4;   // short records: general path, looked at again after a few batches
}

// Fused path: one kernel + summary.
int launch_fused(sk_ctx *c, Slot &s, const BatchArgs &a) {
    cudaStream_t st = a.st;
    if (int rc = next_epoch(c, s, st)) return rc;
    sk::DevInput di[2];
    sk::OutPtrs op;
    make_inputs(c, s, a, di, op);
    s.launches = 0;
    SK_CUDA(cudaEventRecord(s.ev_begin, st));
    int rc;
    const bool two = c->n_inputs == 2;
    if (two) {
        switch (c->fused_ch) {
            case 3: rc = launch_fused_two_ch<3>(c, s, a, di, op); break;
            case 5: rc = launch_fused_two_ch<5>(c, s, a, di, op); break;
            case 9: case 11: rc = launch_fused_two_ch<9>(c, s, a, di, op); break;
            default: rc = launch_fused_two_ch<7>(c, s, a, di, op); break;
        }
    } else {
        switch (c->fused_ch) {
            case 3: rc = launch_fused_ch<3>(c, s, a, di[0], op); break;
            case 5: rc = launch_fused_ch<5>(c, s, a, di[0], op); break;
            case 9: rc = launch_fused_ch<9>(c, s, a, di[0], op); break;
            case 11: rc = launch_fused_ch<11>(c, s, a, di[0], op); break;
            default: rc = launch_fused_ch<7>(c, s, a, di[0], op); break;
        }
    }
    if (rc) return rc;
    SK_CUDA(cudaEventRecord(s.ev_stage[0], st));
    SK_CUDA(cudaEventRecord(s.ev_stage[1], st));
    SK_CUDA(cudaEventRecord(s.ev_stage[2], st));
    if (two) sk::kf2_finalize<<<1, 32, 0, st>>>(s.d_ctl, s.d_res);
    else sk::kf_finalize<<<1, 32, 0, st>>>(di[0], c->dev, s.d_ctl, s.d_res);
    s.launches++;
    SK_CUDA(cudaEventRecord(s.ev_end, st));
    SK_CUDA(cudaGetLastError());
    s.last_fused = true;
    c->n_fused++;
    return SK_OK;
}

// -a N on one input: the single-pass kernel's S1-S6 as an index + verdict pass (kf_fused<CH, 3>: line index and
// {keep, five, kept bases} per record, what K1 and k2_trim_only leave behind, out of one read of the input), then the
// general path's k2_trim_route<true> (reference order, scan), K3 and summary.  The pass gives a batch up exactly
// where the single-pass kernel does; the batch then runs again on K1/K2/K3.
template <int CH>
int launch_index_pass_ch(sk_ctx *c, Slot &s, const BatchArgs &a, const sk::DevInput di[2], const sk::OutPtrs &op) {
    using Cfg = sk::FusedCfg<CH>;
    // two input files: one launch over the tiles of both (tickets dealt in proportion, as in the two-file passes)
    const uint32_t tiles_a = (uint32_t)((a.n[0] + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t tiles_b = c->n_inputs == 2 ? (uint32_t)((a.n[1] + Cfg::kTile - 1) / Cfg::kTile) : 0u;
    const uint32_t tiles = tiles_a + tiles_b;
    const uint32_t stride = c->fused_tiles_cap * sk::kWideStatusStride;
    if (tiles) {
        const int full = c->fused_grid_pass1[(CH - 3) / 2];
        const int grid = tiles < (uint32_t)full ? (int)tiles : full;
        sk::kf_fused<CH, 3><<<grid, sk::kFThreads, Cfg::kSmemPass1, a.st>>>(di[0], c->dev, s.d_ctl, op, s.d_status_f, s.d_status_f + 2 * (size_t)stride, stride,
                                                                         tiles, s.epoch, di[1], tiles_b, nullptr, nullptr,
                                                                         (uint32_t)(c->line_cap / 4 + 1), nullptr, nullptr, s.d_desc[0], nullptr, s.d_desc[1]);
        s.launches++;
        SK_DEBUG_SYNC(a.st, "kf_fused index pass");
    }
    return SK_OK;
}

// -a N (N <= 32), single end, all on the single-pass kernel: the index + verdict pass also leaves every tile's newline
// positions and its kept bytes per queue; one small kernel turns those into every (tile, queue) segment's place in the
// output; the ordered emit pass (kf_fused<CH, 4>) stages a tile queue by queue and flushes up to N segments -- no look-back,
// no K2, no K3.
template <int CH>
int launch_ordered_ch(sk_ctx *c, Slot &s, const BatchArgs &a, const sk::DevInput di[2], const sk::OutPtrs &op) {
    using Cfg = sk::FusedCfg<CH>;
    cudaStream_t st = a.st;
    const uint32_t tiles = (uint32_t)((a.n[0] + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t stride = c->fused_tiles_cap * sk::kWideStatusStride;
    const uint32_t desc_cap = (uint32_t)(c->line_cap / 4 + 1);
    if (tiles) {
        const int full1 = c->fused_grid_pass1[(CH - 3) / 2], full4 = c->fused_grid_ordered[(CH - 3) / 2];
        SK_CUDA(cudaMemsetAsync(s.d_tq + ((size_t)tiles + 1) * 32, 0, ((size_t)tiles / sk::kFTqGroup + 1) * 32 * sizeof(uint32_t), st));   // the group rows
        sk::kf_fused<CH, 3><<<tiles < (uint32_t)full1 ? (int)tiles : full1, sk::kFThreads, Cfg::kSmemPass1, st>>>(
            di[0], c->dev, s.d_ctl, op, s.d_status_f, s.d_status_f + (size_t)stride, stride, tiles, s.epoch, sk::DevInput(), 0u, nullptr, nullptr,
            desc_cap, s.d_nlsave[0], nullptr, s.d_desc[0], s.d_tq);
        SK_DEBUG_SYNC(st, "kf_fused index pass");
        SK_CUDA(cudaEventRecord(s.ev_stage[0], st));
        sk::kfo_offsets<<<1, 1024, 0, st>>>(s.d_ctl, s.d_tq, tiles, c->dev.emu_threads, op.cap[0]);
        SK_DEBUG_SYNC(st, "kfo_offsets");
        SK_CUDA(cudaEventRecord(s.ev_stage[1], st));
        sk::kf_fused<CH, 4><<<tiles < (uint32_t)full4 ? (int)tiles : full4, sk::kFThreads, Cfg::kSmemOrdered, st>>>(
            di[0], c->dev, s.d_ctl, op, s.d_status_f, s.d_status_f + (size_t)stride, stride, tiles, s.epoch, sk::DevInput(), 0u, nullptr, nullptr,
            desc_cap, s.d_nlsave[0], nullptr, s.d_desc[0], s.d_tq);
        SK_DEBUG_SYNC(st, "kf_fused ordered emit");
        s.launches += 3;
    } else {
        SK_CUDA(cudaEventRecord(s.ev_stage[0], st));
        SK_CUDA(cudaEventRecord(s.ev_stage[1], st));
    }
    return SK_OK;
}

int launch_ordered(sk_ctx *c, Slot &s, const BatchArgs &a) {
    cudaStream_t st = a.st;
    if (int rc = next_epoch(c, s, st)) return rc;
    sk::DevInput di[2];
    sk::OutPtrs op;
    make_inputs(c, s, a, di, op);
    s.launches = 0;
    SK_CUDA(cudaEventRecord(s.ev_begin, st));
    int rc;
    switch (c->fused_ch) {
        case 3: rc = launch_ordered_ch<3>(c, s, a, di, op); break;
        case 5: rc = launch_ordered_ch<5>(c, s, a, di, op); break;
        case 9: case 11: rc = launch_ordered_ch<9>(c, s, a, di, op); break;
        default: rc = launch_ordered_ch<7>(c, s, a, di, op); break;
    }
    if (rc) return rc;
    SK_CUDA(cudaEventRecord(s.ev_stage[2], st));
    sk::kf_finalize<<<1, 32, 0, st>>>(di[0], c->dev, s.d_ctl, s.d_res);
    s.launches++;
    SK_CUDA(cudaEventRecord(s.ev_end, st));
    SK_CUDA(cudaGetLastError());
    s.last_fused = true;
    c->n_fused++;
    return SK_OK;
}

int launch_hybrid(sk_ctx *c, Slot &s, const BatchArgs &a) {
    cudaStream_t st = a.st;
    if (int rc = next_epoch(c, s, st)) return rc;
    sk::DevInput di[2];
    sk::OutPtrs op;
    make_inputs(c, s, a, di, op);
    s.launches = 0;
    SK_CUDA(cudaEventRecord(s.ev_begin, st));
    int rc;
    switch (c->fused_ch) {
        case 3: rc = launch_index_pass_ch<3>(c, s, a, di, op); break;
        case 5: rc = launch_index_pass_ch<5>(c, s, a, di, op); break;
        case 9: case 11: rc = launch_index_pass_ch<9>(c, s, a, di, op); break;
        default: rc = launch_index_pass_ch<7>(c, s, a, di, op); break;
    }
    if (rc) return rc;
    SK_CUDA(cudaEventRecord(s.ev_stage[0], st));
    const int resident = c->sm_count * 8;
    const uint64_t max_units = (a.n[0] + a.n[1]) / 4 + 1;
    const uint64_t tiles = (max_units + sk::kK2UnitsPerTile - 1) / sk::kK2UnitsPerTile;
    const int grid = tiles < (uint64_t)resident ? (int)tiles : resident;
    sk::k2_trim_route<true><<<grid, sk::kK2Threads, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1], s.d_status_k2, c->k2_tiles_cap, s.epoch);
    SK_DEBUG_SYNC(st, "k2_trim_route");
    SK_CUDA(cudaEventRecord(s.ev_stage[1], st));
    sk::k3_emit<<<resident, sk::kK3Threads, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, s.d_desc[0], s.d_desc[1], op);
    SK_DEBUG_SYNC(st, "k3_emit");
    SK_CUDA(cudaEventRecord(s.ev_stage[2], st));
    sk::k_finalize<<<1, 32, 0, st>>>(di[0], di[1], c->dev, s.d_ctl, op, s.d_res);
    s.launches += 3;
    SK_CUDA(cudaEventRecord(s.ev_end, st));
    SK_CUDA(cudaGetLastError());
    s.last_fused = true;
    c->n_fused++;
    return SK_OK;
}

int launch_batch(sk_ctx *c, Slot &s, const BatchArgs &a) {
    s.last = a;
    // (the back-off is counted down where results are read, rerun_if_needed: a caller that queues batches
    //  without looking at their summaries never learns that a batch failed, so it must not drift back either)
    if (c->fused_eligible && c->fused_backoff == 0) return launch_fused(c, s, a);
    if (c->ordered_eligible && c->fused_backoff == 0) return launch_ordered(c, s, a);
    if (c->hybrid_eligible && c->fused_backoff == 0) return launch_hybrid(c, s, a);
    return launch_general(c, s, a);
}

// The summary of the slot's batch is in h_res.  If the fused kernel gave up, run the batch again on
// the general path (same stream) and wait for its summary.
int rerun_if_needed(sk_ctx *c, Slot &s) {
    if (!s.last_fused || !(s.h_res->index_overflow & 4u)) {
        if (!s.last_fused && c->fused_backoff > 0) c->fused_backoff--;
        if (!s.last_fused) {
            const uint64_t nrec = s.h_res->records[0] + s.h_res->records[1];
            if (nrec) c->long_records = (s.h_res->consumed[0] + s.h_res->consumed[1]) / nrec >= sk::kK2LongRecordBytes;
        }
        if (s.last_fused && s.h_res->err_kind == 0 && !(s.h_res->index_overflow & 3u)) adapt_fused(c, *s.h_res, false);
        return SK_OK;
    }
    // The single-pass kernel gave the batch up (a record longer than a tile's halo, a data error, ...): this batch
    // runs again on the general path.  A first failure costs nothing more (one long read in a file of short ones
    // must not push the following batches onto the slower path); a second one in a row sends the next 16 batches
    // to the general path, then 32, 64, ... up to 4096 (a file of long reads pays one wasted -- and, since the
    // kernel stops working once a tile has failed, short -- launch in thousands of batches, not one in 17).
    c->fused_backoff = c->fused_fail_streak == 0 ? 0 : 16 << (c->fused_fail_streak < 9 ? c->fused_fail_streak - 1 : 8);
    c->fused_fail_streak++;
    adapt_fused(c, *s.h_res, true);
    c->n_rerun++;
    if (int rc = launch_general(c, s, s.last)) return rc;
    SK_CUDA(cudaMemcpyAsync(s.h_res, s.d_res, sizeof(sk::DevResult), cudaMemcpyDeviceToHost, s.last.st));
    SK_CUDA(cudaStreamSynchronize(s.last.st));
    return SK_OK;
}

void fill_result(const Slot &s, sk_result *res) {
    const sk::DevResult &r = *s.h_res;
    memset(res, 0, sizeof *res);
    for (int k = 0; k < 3; ++k) res->out_bytes[k] = r.out_bytes[k];
    for (int k = 0; k < 2; ++k) {
        // device offsets are relative to the aligned-down batch pointer; consumed counts from `start`
        const uint64_t first = s.first[k];
        res->consumed[k] = r.consumed[k] > first ? r.consumed[k] - first : 0;
        res->records[k] = r.records[k];
    }
    res->kept = r.counters[0]; res->discard = r.counters[1];
    res->kept_p = r.counters[2]; res->discard_p = r.counters[3];
    res->kept_s1 = r.counters[4]; res->kept_s2 = r.counters[5];
    res->discard_s1 = r.counters[6]; res->discard_s2 = r.counters[7];
    res->error.kind = r.err_kind; res->error.file = r.err_file; res->error.record = r.err_record;
    res->error.position = r.err_position; res->error.byte = r.err_byte;
    for (int k = 0; k < 4; ++k) {
        res->error.line_off[k] = r.err_line_off[k] + s.base[r.err_file & 1];
        res->error.line_len[k] = r.err_line_len[k];
    }
    res->kernel_launches = s.launches;
    res->fused = s.last_fused ? 1u : 0u;
}

int stage_times(const Slot &s, sk_result *res) {
    SK_CUDA(cudaEventElapsedTime(&res->kernel_ms, s.ev_begin, s.ev_end));
    SK_CUDA(cudaEventElapsedTime(&res->stage_ms[0], s.ev_begin, s.ev_stage[0]));
    SK_CUDA(cudaEventElapsedTime(&res->stage_ms[1], s.ev_stage[0], s.ev_stage[1]));
    SK_CUDA(cudaEventElapsedTime(&res->stage_ms[2], s.ev_stage[1], s.ev_stage[2]));
    SK_CUDA(cudaEventElapsedTime(&res->stage_ms[3], s.ev_stage[2], s.ev_end));
    return SK_OK;
}

int check_capacity(const sk_ctx *c, const Slot &s) {
    if (s.h_res->index_overflow & 1u) {
        set_err("batch has more than %u lines per input; lower SICKLE_B200_MIN_LINE_BYTES or shrink the batch", c->line_cap);
        return SK_E_CAPACITY;
    }
    if (s.h_res->index_overflow & 2u) { set_err("output stream larger than its buffer"); return SK_E_CAPACITY; }
    return SK_OK;
}

}  // namespace

extern "C" {

int sk_abi_version(void) { return SK_ABI_VERSION; }

int sk_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { set_err("cudaGetDeviceCount: %s", cudaGetErrorString(e)); return -1; }
    return n;
}

const char *sk_last_error(void) { return g_err; }

uint64_t sk_slot_bytes(const sk_ctx *ctx) { return ctx ? ctx->slot_bytes : 0; }

sk_ctx *sk_create(int device, uint64_t slot_bytes, int n_slots, const sk_params *params) {
    if (!params || slot_bytes == 0 || slot_bytes > kMaxSlotBytes || n_slots < 0 || n_slots > 64) {
        set_err("sk_create: bad arguments (slot_bytes=%llu n_slots=%d)", (unsigned long long)slot_bytes, n_slots);
        return nullptr;
    }
    sk::DevParams dp;
    if (make_dev_params(*params, dp) != SK_OK) return nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        set_err("sk_create: CUDA device %d not available (%d devices) -- there is no CPU fallback", device, ndev);
        return nullptr;
    }
    // SICKLE_B200_DEBUG_INIT=1: where the start-up time goes (CUDA context vs. pinned / device allocations)
    const bool dbg_init = getenv("SICKLE_B200_DEBUG_INIT") != nullptr;
    const auto t_init0 = std::chrono::steady_clock::now();
    if (cudaSetDevice(device) != cudaSuccess) { set_err("cudaSetDevice(%d) failed", device); return nullptr; }
    if (dbg_init) cudaFree(nullptr);   // forces the context into existence here, so that the two figures separate
    const auto t_init1 = std::chrono::steady_clock::now();
    sk_ctx *c = new (std::nothrow) sk_ctx();
    if (!c) { set_err("out of memory"); return nullptr; }
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    c->slot_bytes = (slot_bytes + 15) & ~15ull;
    c->params = *params;
    c->dev = dp;
    c->n_inputs = params->mode == SK_MODE_PE_2FILE ? 2 : 1;
    // line index capacity: average line >= SICKLE_B200_MIN_LINE_BYTES bytes (default 4, '\n' included)
    uint64_t min_line = 4;
    if (const char *e = getenv("SICKLE_B200_MIN_LINE_BYTES")) { long v = atol(e); if (v >= 1 && v <= 1024) min_line = (uint64_t)v; }
    c->line_cap = (uint32_t)(c->slot_bytes / min_line + 8) & ~3u;
    c->k1_tiles_cap = (uint32_t)((c->slot_bytes + sk::kK1TileBytes - 1) / sk::kK1TileBytes) + 1;
    c->k2_tiles_cap = (uint32_t)(((uint64_t)c->line_cap / 4 + sk::kK2UnitsPerTile - 1) / sk::kK2UnitsPerTile) + 2;
    c->fused_tiles_cap = (uint32_t)(c->slot_bytes / kFusedMinTile) + 2;
    // SICKLE_B200_PATH = auto (default) | general | fused ; SICKLE_B200_FUSED_CH = 5 | 7 | 9 | 11
    c->fused_eligible = dp.emu_threads == 1;
    c->hybrid_eligible = dp.emu_threads > 1;
    c->ordered_eligible = c->hybrid_eligible && dp.emu_threads <= 32 && params->mode == SK_MODE_SE;
    if (const char *e = getenv("SICKLE_B200_ORDERED")) { if (atoi(e) == 0) c->ordered_eligible = false; }
    c->verdict_cap = (uint32_t)(c->slot_bytes / 32 + 64);   // records of 32 bytes and more (shorter ones: general path)
    if (const char *e = getenv("SICKLE_B200_K2_SPLIT")) c->k2_split = atoi(e) != 0;
    if (const char *e = getenv("SICKLE_B200_PATH")) { if (!strcmp(e, "general")) c->fused_eligible = c->hybrid_eligible = c->ordered_eligible = false; }
    if (const char *e = getenv("SICKLE_B200_FUSED_CH")) { c->fused_ch = atoi(e); c->fused_ch_fixed = true; }
    if ((c->fused_eligible || c->hybrid_eligible) && setup_fused(c) != SK_OK) { delete c; return nullptr; }
    c->host_buffers = n_slots > 0;
    const int ns = n_slots > 0 ? n_slots : 1;
    c->slots.resize(ns);
    for (int i = 0; i < ns; ++i) {
        if (alloc_slot(c, c->slots[i], c->host_buffers) != SK_OK) {
            for (auto &s : c->slots) free_slot(s);
            delete c;
            return nullptr;
        }
    }
    if (dbg_init) {
        const auto t_init2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[sickle_b200] sk_create: CUDA context %.3f s, %d slot(s) of %llu bytes (pinned + device buffers, kernels' attributes) %.3f s\n",
                std::chrono::duration<double>(t_init1 - t_init0).count(), ns, (unsigned long long)c->slot_bytes,
                std::chrono::duration<double>(t_init2 - t_init1).count());
    }
    return c;
}

void sk_destroy(sk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        free_slot(s);
    }
    delete ctx;
}

char *sk_in_buffer(sk_ctx *ctx, int slot, int which) {
    if (!ctx || !ctx->host_buffers || slot < 0 || slot >= (int)ctx->slots.size() || which < 0 || which >= ctx->n_inputs) {
        set_err("sk_in_buffer: bad slot/which");
        return nullptr;
    }
    return ctx->slots[slot].h_in[which];
}

int sk_upload(sk_ctx *ctx, int slot, int which, uint64_t offset, uint64_t nbytes) {
    if (!ctx || !ctx->host_buffers || slot < 0 || slot >= (int)ctx->slots.size() || which < 0 || which >= ctx->n_inputs) {
        set_err("sk_upload: bad slot/which");
        return SK_E_ARG;
    }
    Slot &s = ctx->slots[slot];
    if (s.busy) { set_err("sk_upload: slot %d still busy (call sk_wait first)", slot); return SK_E_ARG; }
    if (offset + nbytes > ctx->slot_bytes || s.up_hi[which] != s.up_lo[which]) {
        set_err("sk_upload: range outside the slot, or a range was already uploaded for this batch");
        return SK_E_ARG;
    }
    if (!nbytes) return SK_OK;
    SK_CUDA(cudaSetDevice(ctx->device));
    SK_CUDA(cudaMemcpyAsync(s.d_in[which] + offset, s.h_in[which] + offset, nbytes, cudaMemcpyHostToDevice, s.stream));
    s.up_lo[which] = offset;
    s.up_hi[which] = offset + nbytes;
    return SK_OK;
}

int sk_submit(sk_ctx *ctx, int slot, uint64_t start0, uint64_t end0, uint64_t start1, uint64_t end1) {
    if (!ctx || !ctx->host_buffers || slot < 0 || slot >= (int)ctx->slots.size()) { set_err("sk_submit: bad slot"); return SK_E_ARG; }
    const uint64_t st[2] = {start0, start1}, en[2] = {end0, end1};
    for (int i = 0; i < 2; ++i) {
        if (st[i] > en[i] || en[i] > ctx->slot_bytes || (i >= ctx->n_inputs && en[i] != st[i])) {
            set_err("sk_submit: bad byte range [%llu, %llu) for input %d (slot_bytes %llu)", (unsigned long long)st[i],
                    (unsigned long long)en[i], i, (unsigned long long)ctx->slot_bytes);
            return SK_E_ARG;
        }
    }
    Slot &s = ctx->slots[slot];
    if (s.busy) { set_err("sk_submit: slot %d still busy (call sk_wait first)", slot); return SK_E_ARG; }
    SK_CUDA(cudaSetDevice(ctx->device));
    BatchArgs a;
    for (int i = 0; i < ctx->n_inputs; ++i) {
        // upload what sk_upload has not covered: [start, min(end, up_lo)) and [max(start, up_hi), end)
        uint64_t lo = s.up_lo[i], hi = s.up_hi[i];
        if (hi == lo) { lo = hi = st[i]; }
        if (lo < st[i]) lo = st[i];
        if (hi > en[i]) hi = en[i];
        if (hi < lo) hi = lo;
        if (lo > st[i])
            SK_CUDA(cudaMemcpyAsync(s.d_in[i] + st[i], s.h_in[i] + st[i], lo - st[i], cudaMemcpyHostToDevice, s.stream));
        if (en[i] > hi)
            SK_CUDA(cudaMemcpyAsync(s.d_in[i] + hi, s.h_in[i] + hi, en[i] - hi, cudaMemcpyHostToDevice, s.stream));
        s.up_lo[i] = s.up_hi[i] = 0;
        s.base[i] = st[i] & ~15ull;
        s.first[i] = (uint32_t)(st[i] - s.base[i]);
        a.in[i] = s.d_in[i] + s.base[i];
        a.first[i] = s.first[i];
        a.n[i] = en[i] - s.base[i];
    }
    for (int k = 0; k < 3; ++k) { a.out[k] = s.d_out[k]; a.cap[k] = s.out_cap[k]; }
    a.st = s.stream;
    int rc = launch_batch(ctx, s, a);
    if (rc != SK_OK) return rc;
    SK_CUDA(cudaMemcpyAsync(s.h_res, s.d_res, sizeof(sk::DevResult), cudaMemcpyDeviceToHost, s.stream));
    SK_CUDA(cudaEventRecord(s.ev_done, s.stream));
    s.busy = true;
    s.device_mode = false;
    return SK_OK;
}

int sk_wait(sk_ctx *ctx, int slot, sk_result *res) {
    if (!ctx || !res || slot < 0 || slot >= (int)ctx->slots.size()) { set_err("sk_wait: bad arguments"); return SK_E_ARG; }
    Slot &s = ctx->slots[slot];
    if (!s.busy || s.device_mode) { set_err("sk_wait: slot %d has no submitted batch", slot); return SK_E_ARG; }
    SK_CUDA(cudaSetDevice(ctx->device));
    SK_CUDA(cudaEventSynchronize(s.ev_done));
    s.busy = false;
    if (int rc = rerun_if_needed(ctx, s)) return rc;
    fill_result(s, res);
    if (int rc = check_capacity(ctx, s)) return rc;
    if (int rc = stage_times(s, res)) return rc;
    if (res->error.kind == 0) {
        for (int k = 0; k < 3; ++k) {
            if (res->out_bytes[k] && s.h_out[k])
                SK_CUDA(cudaMemcpyAsync(s.h_out[k], s.d_out[k], res->out_bytes[k], cudaMemcpyDeviceToHost, s.stream));
            res->out[k] = s.h_out[k];
        }
        SK_CUDA(cudaStreamSynchronize(s.stream));
    } else {
        for (int k = 0; k < 3; ++k) res->out_bytes[k] = 0;
    }
    return SK_OK;
}

int sk_trim_device(sk_ctx *ctx, int slot, const void *in0, uint64_t n0, const void *in1, uint64_t n1,
                   void *const out[3], const uint64_t out_cap[3], void *stream) {
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size() || !out || !out_cap) { set_err("sk_trim_device: bad arguments"); return SK_E_ARG; }
    if (n0 > ctx->slot_bytes || n1 > ctx->slot_bytes || (ctx->n_inputs == 1 && n1) ||
        (reinterpret_cast<uintptr_t>(in0) & 15) || (reinterpret_cast<uintptr_t>(in1) & 15)) {
        set_err("sk_trim_device: inputs must be 16-byte aligned and <= slot_bytes");
        return SK_E_ARG;
    }
    Slot &s = ctx->slots[slot];
    if (s.busy && !s.device_mode) { set_err("sk_trim_device: slot %d busy", slot); return SK_E_ARG; }
    SK_CUDA(cudaSetDevice(ctx->device));
    BatchArgs a;
    a.in[0] = (const uint8_t *)in0; a.in[1] = (const uint8_t *)in1;
    a.n[0] = n0; a.n[1] = n1;
    for (int k = 0; k < 3; ++k) { a.out[k] = (uint8_t *)out[k]; a.cap[k] = out_cap[k]; }
    a.st = stream ? (cudaStream_t)stream : s.stream;
    s.base[0] = s.base[1] = 0;
    s.first[0] = s.first[1] = 0;
    int rc = launch_batch(ctx, s, a);
    if (rc != SK_OK) return rc;
    s.busy = true;
    s.device_mode = true;
    return SK_OK;
}

int sk_result_device(sk_ctx *ctx, int slot, void *stream, sk_result *res) {
    if (!ctx || !res || slot < 0 || slot >= (int)ctx->slots.size()) { set_err("sk_result_device: bad arguments"); return SK_E_ARG; }
    Slot &s = ctx->slots[slot];
    if (!s.busy || !s.device_mode) { set_err("sk_result_device: nothing submitted on slot %d", slot); return SK_E_ARG; }
    SK_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = s.last.st;
    (void)stream;
    SK_CUDA(cudaMemcpyAsync(s.h_res, s.d_res, sizeof(sk::DevResult), cudaMemcpyDeviceToHost, st));
    SK_CUDA(cudaStreamSynchronize(st));
    s.busy = false;
    if (int rc = rerun_if_needed(ctx, s)) return rc;
    fill_result(s, res);
    if (int rc = stage_times(s, res)) return rc;
    return check_capacity(ctx, s);
}

#ifdef SK_PHASE_TIMING
// debug build only: cycles per fused-kernel phase, summed over tiles (thread 0 of every CTA)
int sk_debug_phase_cycles(unsigned long long out[12], int reset) {
    SK_CUDA(cudaDeviceSynchronize());
    SK_CUDA(cudaMemcpyFromSymbol(out, sk::g_phase_cycles, sizeof(unsigned long long) * 12));
    unsigned long long w[4];
    SK_CUDA(cudaMemcpyFromSymbol(w, sk::g_walk_dbg, sizeof w));
    fprintf(stderr, "[walk dbg] walks %llu steps %llu threads-that-spun %llu spin-iterations %llu\n", w[0], w[1], w[2], w[3]);
    if (reset) {
        unsigned long long z[12] = {0};
        SK_CUDA(cudaMemcpyToSymbol(sk::g_phase_cycles, z, sizeof z));
        SK_CUDA(cudaMemcpyToSymbol(sk::g_walk_dbg, z, sizeof(unsigned long long) * 4));
    }
    return SK_OK;
}
#endif

}  // extern "C"
