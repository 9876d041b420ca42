// kf_fused.cuh -- single-pass fused kernel for short records: parse + trim + route + emit.
//
// One pass over the batch: every input byte is read from HBM once and every output byte written
// once (the general K1/K2/K3 path reads the input three times and round-trips a line index and a
// descriptor table).  Same results as K1+K2+K3 -- it covers, per tile,
//   * the Batch/FQEntry line split              (reference src/GZReader.cpp:76-92, src/FQEntry.cpp:8-18)
//   * FQEntry::validate                         (src/FQEntry.cpp:53-97)
//   * sliding_window / get_quality_num          (src/trim.cpp:3-140)
//   * keep / singles / discard routing          (src/trim_single.cpp:382-404, src/trim_paired.cpp:531-567)
//   * record formatting                         (src/trim_single.cpp:393-396, src/trim_paired.cpp:506-513)
// for SK_MODE_SE, SK_MODE_PE_INTER and SK_MODE_PE_INTER_M in input order (emulate_threads == 1); two input files and
// the reference's -a N order run the kernel twice (template parameter PASS, below).
//
// Tile structure (256 threads, persistent CTAs, tiles handed out by an atomic ticket):
//   S1  region = tile (224 threads x CH x 16 B) + halo (32 threads x CH x 16 B) -> shared memory by
//       one TMA bulk copy (cp.async.bulk + mbarrier).  The halo lets a record that starts in the tile
//       finish.  The tile one grid-width ahead is prefetched into L2.
//   S2  each thread scans its CH*16 contiguous bytes (odd CH => conflict-free LDS.128): SWAR newline
//       test (3 ALU ops / word) -> bit masks -> popc.
//   S3  block scan -> rank of every newline; positions to shared memory (u16).
//   S4  decoupled look-back #1 over the tiles' newline counts -> global line number of the tile.
//   S5  records are lines 4r..4r+3; a record (pair) belongs to the tile holding the newline in
//       front of it.
//   S6  validate + sliding window out of shared memory, one or two lanes per read and branch-free:
//       32 windows per step, window totals by dp4a straight from the packed quality words, the sign
//       of every total shifted into a bit mask, a cheap range screen per word; resolved once per step.
//   S7  routing, 2-stream block scan; the tile's output sizes are published for look-back #2.
//   S8a two lanes per record copy the trimmed record into a shared staging buffer.
//   S8b one tile LATER, by the "flush group" -- the 3 or 4 warps that never own a record -- while the
//       record warps run S5-S7 of the next tile: look-back #2 -> byte offsets in the output streams,
//       then the staged bytes go out with destination-aligned 16-byte stores (128-bit funnel shift
//       for the destination's phase).
// Two input files (SK_MODE_PE_2FILE: mate 1 in one file, mate 2 in the other, paired by record number,
// reference src/trim_paired.cpp:328-338,483-567) run the same kernel twice, template parameter PASS:
//   PASS 1  tiles of both files (tickets dealt in proportion to the files' sizes), S1-S6 only: every complete
//           record's verdict {keep, five, kept bases, bytes of name + line 3} goes into an 8-byte entry of a
//           per-file table indexed by record number; the files' record counts are summed up.
//   (kf2_between: units = min of the two record counts; ticket counter back to zero)
//           PASS 1 also leaves every tile's newline positions and line numbers in global memory (kFNlSlot bytes).
//   PASS 2  the same tiles again: S1 fetches the tile and, with a second bulk copy on the same mbarrier, what PASS 1
//           saved for it (no S2-S4: no masks, no scan, no look-back #1), S5, then instead of trimming, a record looks
//           up its own entry and its mate's, is routed (both kept -> its file's stream; one kept -> singles), and S7-S8 run as for
//           interleaved pairs: stream 0 of a file-f tile is output stream f, flushed flat; the singles stream
//           is laid out in pair order, so a tile also reserves room for the singles its mates' tiles write
//           (sizes from the mates' entries) and copies its own singles out one by one.
// The input is read twice (650 + 280 bytes per read instead of 325 + 280), which a kernel at 30 % of the
// HBM roofline can afford; in exchange no tile ever waits for another tile's trimming.
// PASS 0 is the single pass described above.
// The reference's -a N output order (emulate_threads = N > 1: record k of a batch is dealt to queue (k+1) % N and the
// queues are written one after the other, src/trim_single.cpp:263,273-274,374-428) is two passes as well:
//   PASS 3  index + verdict pass: S1-S6 over one input or two; per record {keep, five, kept bases} into the general
//           path's descriptor table.  Either (tq == nullptr) the line index K1 would have written goes to line_end and
//           k2_trim_route<true> + k3_emit finish the batch (any mode, any N), or (single end, N <= 32) the tile's newline
//           positions are saved as in PASS 1 and its kept bytes per queue are counted into tq;
//   (kfo_offsets: every (tile, queue) segment's place in the output)
//   PASS 4  ordered emit: tile + saved positions by TMA, verdicts from the table, the tile staged queue by queue and
//           flushed as up to N segments -- no look-back in this pass.
//
// CH (3, 5, 7 or 9: 11 / 18 / 25 / 32 KB tiles) is chosen by the host per batch so that a tile holds at most
// ~112 records.  Anything this kernel cannot handle exactly -- a record longer than the halo, more
// than 128 records or 1024 newlines in a tile, or ANY data error -- sets Control::fast_fail; the host
// then re-runs the batch through the general path, which also produces the reference's error details.
#pragma once

#include "k1_index.cuh"
#include "k2_trim.cuh"
#include "sk_copy.cuh"
#include "sk_device.cuh"
#include "trim_lane.cuh"

namespace sk {

// Debug only (-DSK_PHASE_TIMING): barrier-to-barrier cycles of thread 0, summed over tiles.
#ifdef SK_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[12];
#define SK_TICK(k)                                                            \
    do {                                                                      \
        if (tid == 0) {                                                       \
            const long long t_now = clock64();                                \
            atomicAdd(&g_phase_cycles[k], (unsigned long long)(t_now - t_prev)); \
            t_prev = t_now;                                                   \
        }                                                                     \
    } while (0)
#else
#define SK_TICK(k) do { } while (0)
#endif

// Timing-only knock-outs (-DSK_KO_S6 / SK_KO_S8A / SK_KO_FLUSH / SK_KO_LB1 / SK_KO_LB2, never in a shipped
// build: the output is WRONG): each removes one phase so that `bench.py --kernel-only` shows what that
// phase costs in throughput rather than in instruction share (profiles/ab_variants.sh).  SK_KO_LB1
// assumes the bench's fixed 325-byte records.
#ifndef SK_MAX_CTAS
#define SK_MAX_CTAS 4      // resident CTAs per SM the register budget is set for (__launch_bounds__)
#endif
constexpr int kFThreads = 256;
constexpr int kFTileThreads = 224;   // 7 warps own tile bytes, the 8th warp scans the halo
constexpr int kFMaxNl = 1024;        // newline positions per region (2 KB; also holds 128 record descriptors)
// Two files: PASS 1 leaves every tile's newline positions and {global line number, newlines in the tile, newlines
// in the region} in global memory, kFNlSlot bytes per tile, and PASS 2 fetches them with the tile's bytes instead of
// finding the newlines again (no masks, no scan, no position stores, no look-back #1).
constexpr int kFNlSlot = kFMaxNl * 2 + 16;
// -a N: d_tq holds one row of 32 counters per tile, one row of queue bases, and one row per group of kFTqGroup tiles
constexpr uint32_t kFTqGroup = 16;
constexpr uint32_t kTicketPoison = 0x40000000u;   // or-ed into the ticket counter by a tile that gives the batch up (tiles < 2^30)

template <int CH>
struct FusedCfg {
    static_assert(CH % 2 == 1, "CH must be odd (conflict-free 16-byte shared loads at stride CH*16)");
    static constexpr int kBytesPerThread = CH * 16;
    static constexpr int kRegion = kFThreads * kBytesPerThread;
    static constexpr int kTile = kFTileThreads * kBytesPerThread;
    static constexpr int kInBytes = kRegion + 96;        // the window loop reads up to ~50 bytes past a record
    static constexpr int kOutBytes = kTile + 768 + 96;   // a tile of output + the last record's overhang + phase shifts
    // CH = 7: 56,768 B -> four CTAs per SM (4 x (56,768 + 248 static + 1,024 reserved) <= 232,448)
    static constexpr size_t kSmem = (size_t)kInBytes + kOutBytes + kFMaxNl * 2;
    static constexpr size_t kSmemTwoFile = kSmem + 16 + 128 * 8;   // PASS 2: + the tile's saved line numbers, + one (offset, length) pair per record: its own singles
    // PASS 4: + the tile's saved line numbers, + bytes per record (128 x 4), queue totals (32 x 4), and the staged
    // tile's segments for the deferred flush: start, length (32 x 4 each), destination (32 x 8)
    static constexpr size_t kSmemOrdered = kSmem + 16 + 512 + 128 + 128 + 128 + 256;
    static constexpr size_t kSmemPass1 = (size_t)kInBytes + kFMaxNl * 2;   // PASS 1 stages nothing: five CTAs per SM
    // (the passes that stage nothing compile to 48 registers without a spill: five CTAs per SM instead of four,
    //  -3 % on two files and on -a 8, profiles/r2_call25.log)
#ifndef SK_MAX_CTAS_PASS1
#define SK_MAX_CTAS_PASS1 5
#endif
    static constexpr int kCtasPerSmPass1 = (int)(232448 / (kSmemPass1 + 1024 + 256)) > SK_MAX_CTAS_PASS1 ? SK_MAX_CTAS_PASS1 : (int)(232448 / (kSmemPass1 + 1024 + 256));
    // CTAs per SM by shared memory (232,448 B per SM, 1,024 B reserved per CTA); also the register budget
    static constexpr int kCtasPerSm = (int)(232448 / (kSmem + 1024 + 256)) > SK_MAX_CTAS ? SK_MAX_CTAS : (int)(232448 / (kSmem + 1024 + 256));
};


// The deferred half of a tile (S7b + S8b), run by the "flush group" -- warps 4..7, which never own a
// record -- while warps 0..3 validate and trim the next tile: look-back #2 over the output sizes,
// then the staged bytes go out.  gtid = thread index inside the group (0..127), named barrier 1.
constexpr int kFlushBarrier = 1;
// Two files (PASS 2): `main` = the output stream of the tile's file (0 or 1), `last` = the tile is its file's last
// one, and the singles are not flushed flat -- the singles area has holes where the mates' tiles write -- but
// copied record by record from the list (offset inside the tile's singles range, length; length 0 = not mine).
__device__ __forceinline__ void flush_previous_tile(unsigned long long *const st_out[2], uint32_t p_tile, uint32_t p_tot0,
                                                    uint32_t p_tot1, int nstreams, uint32_t epoch, int gtid, int gwarps,
                                                    unsigned long long (*s_lb)[2], Control *__restrict__ ctl,
                                                    const OutPtrs &outs, const uint8_t *__restrict__ s_out, bool last,
                                                    int main = 0, const uint2 *__restrict__ singles = nullptr, uint32_t nsingles = 0) {
    const unsigned long long agg[2] = {p_tot0, p_tot1};
    unsigned long long ex[2];
#ifdef SK_PHASE_TIMING
    const long long t_in = clock64();
#endif
#ifdef SK_KO_LB2
    ex[0] = (unsigned long long)p_tile * 30000ull; ex[1] = (unsigned long long)p_tile * 3000ull;   // no walk: made-up offsets
#else
    block_walk(st_out, p_tile, agg, nstreams, epoch, gtid, s_lb, ex, gwarps, kFlushBarrier);
#endif
#ifdef SK_PHASE_TIMING
    if (gtid == 0) atomicAdd(&g_phase_cycles[5], (unsigned long long)(clock64() - t_in));   // look-back #2 (flush group)
#endif
    if (gtid == 0 && last) {
        ctl->out_bytes[main] = ex[0] + agg[0];
        if (main == 0) ctl->out_bytes[2] = ex[1] + agg[1];   // (both files' singles chains end at the same total)
    }
    const bool cap_ok = ex[0] + p_tot0 <= outs.cap[main] && (p_tot1 == 0 || (outs.p[2] && ex[1] + p_tot1 <= outs.cap[2]));
    if (!cap_ok) {
        if (gtid == 0) ctl->index_overflow = 2u;
        return;
    }
#if !defined(SK_KO_FLUSH)
    flush_realigned(outs.p[main] + ex[0], s_out, 0u, p_tot0, gtid, gwarps * 32);
    if (singles == nullptr) {
        if (p_tot1) flush_realigned(outs.p[2] + ex[1], s_out, (p_tot0 + 15u) & ~15u, p_tot1, gtid, gwarps * 32);
    } else if (p_tot1) {
        // own singles, one thread per record (few per tile): staging -> global, both at arbitrary phases
        uint8_t *const g = outs.p[2] + ex[1];
        const uint32_t ph = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15u);
        const uint32_t base1 = (p_tot0 + 15u) & ~15u;
        for (uint32_t i = (uint32_t)gtid; i < nsingles; i += (uint32_t)gwarps * 32u) {
            const uint2 e = singles[i];
            if (e.y) smem_copy(g - ph, ph + e.x, s_out, base1 + e.x, e.y);
        }
    }
#endif
#ifdef SK_PHASE_TIMING
    if (gtid == 0) atomicAdd(&g_phase_cycles[7], (unsigned long long)(clock64() - t_in));   // look-back #2 + flush (flush group)
#endif
}

// PASS 4 (-a N order): the staged tile is up to 32 segments, one per queue of the reference's round-robin dealing,
// each with its own place in the output (known before the pass starts: no look-back).  Warp gw of the flush group
// takes the queues gw, gw + 4, ...
__device__ __forceinline__ void flush_ordered_tile(const OutPtrs &outs, const uint8_t *__restrict__ s_out, const uint32_t *__restrict__ seg,
                                                   const uint32_t *__restrict__ len, const unsigned long long *__restrict__ dst, int nq, int gtid) {
    const int gw = gtid >> 5, gl = gtid & 31;
    for (int q = gw; q < nq; q += 4) {
        const uint32_t n = len[q];
        if (n) flush_realigned(outs.p[0] + dst[q], s_out, seg[q], n, gl, 32);
    }
}

// Software-pipelined over tiles: the output of tile t is staged in shared memory right after it is
// trimmed, its size is published, and only one tile later -- after the next tile has been loaded and
// scanned -- does the CTA ask for tile t's output offset and flush.  By then every predecessor has
// long published its size, so the second look-back does not wait; with hundreds of heavy tiles in
// flight an immediate look-back makes every tile wait for the slowest predecessor (measured: 22 % of
// the tile time).
// Entry of the two-file verdict table: [15:0] bytes of name + line 3 + 4 newlines, [31:16] kept bases,
// [47:32] first kept base, [48] keep.
__device__ __forceinline__ unsigned long long pack_verdict(bool keep, uint32_t five, uint32_t nkeep, uint32_t fixed) {
    return (unsigned long long)(fixed & 0xffffu) | ((unsigned long long)(nkeep & 0xffffu) << 16) |
           ((unsigned long long)(five & 0xffffu) << 32) | ((unsigned long long)(keep ? 1u : 0u) << 48);
}

template <int CH, int PASS = 0>
__global__ void __launch_bounds__(kFThreads, (PASS == 1 || PASS == 3) ? FusedCfg<CH>::kCtasPerSmPass1 : FusedCfg<CH>::kCtasPerSm)
kf_fused(DevInput in_a, DevParams P, Control *__restrict__ ctl, OutPtrs outs,
         unsigned long long *__restrict__ status_nl_all, unsigned long long *__restrict__ status_out_all /* [2][stride] per file */,
         uint32_t status_stride, uint32_t num_tiles, uint32_t epoch,
         // two files only (PASS 1, 2): the second input, its share of the tiles, the verdict tables
         DevInput in_b = DevInput(), uint32_t tiles_b = 0, unsigned long long *__restrict__ tab_a = nullptr,
         unsigned long long *__restrict__ tab_b = nullptr, uint32_t tab_cap = 0,
         uint8_t *__restrict__ nlsave_a = nullptr, uint8_t *__restrict__ nlsave_b = nullptr /* kFNlSlot bytes per tile */,
         // PASS 3 only: the general path's descriptor table (tab_cap entries); the line index goes to in_a.line_end
         RecDesc *__restrict__ desc_out = nullptr,
         // PASS 3 (optional) / PASS 4: [num_tiles + 1][32] kept bytes per tile and queue (record k of the batch is dealt to
         // queue (k+1) % N, src/trim_single.cpp:263,273-274); kfo_offsets turns the counters into output offsets
         uint32_t *__restrict__ tq = nullptr,
         // PASS 3 on two input files (tiles_b > 0): the second file's descriptor table; its line index goes to in_b.line_end
         RecDesc *__restrict__ desc_out_b = nullptr) {
    using Cfg = FusedCfg<CH>;
    constexpr bool kNoEmit = PASS == 1 || PASS == 3;   // these passes stage and write no records
    constexpr bool kSaved = PASS == 2 || PASS == 4;    // newline positions and line numbers come from the pass before
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t *s_in = smem;
    uint8_t *s_out = smem + Cfg::kInBytes;
    uint16_t *s_nl = reinterpret_cast<uint16_t *>(smem + Cfg::kInBytes + (kNoEmit ? 0 : Cfg::kOutBytes));
    uint4 *s_desc = reinterpret_cast<uint4 *>(s_nl);   // S7/S8a record descriptors alias the newline positions
    const uint4 *s_meta = reinterpret_cast<const uint4 *>(smem + Cfg::kSmem);   // PASS 2, 4: lands right behind s_nl
    uint2 *s_single = reinterpret_cast<uint2 *>(smem + Cfg::kSmem + 16);   // PASS 2 only (Cfg::kSmemTwoFile)
    uint32_t *s_rbytes = reinterpret_cast<uint32_t *>(smem + Cfg::kSmem + 16);            // PASS 4 only (Cfg::kSmemOrdered)
    uint32_t *s_qacc = s_rbytes + 128, *s_fseg = s_qacc + 32, *s_flen = s_fseg + 32;
    unsigned long long *s_fdst = reinterpret_cast<unsigned long long *>(s_flen + 32);
    __shared__ uint32_t s_q[32];   // PASS 3: kept bytes of the tile per queue
    __shared__ uint32_t s_tile;
    __shared__ uint32_t warp_tot[kFThreads / 32];
    __shared__ uint32_t warp_tot2[kFThreads / 32][2];
    __shared__ unsigned long long s_lb[kFThreads / 32][2];   // look-back scratch
    __shared__ uint32_t s_fail;
    // Per-warp running totals of the batch summary (records, end of the last complete record, the reference's
    // counters).  They go to the Control block once, when the CTA runs out of tiles: one atomic per counter and
    // warp and TILE -- 10,000 tiles x 4 warps x up to 8 counters, all on one or two cache lines -- kept a single
    // L2 slice busy for a good part of the launch (paired end: +37 % kernel time over single end).
    __shared__ uint32_t s_acc[kFThreads / 32][8];
    __shared__ __align__(8) unsigned long long s_mbar;   // completion of the S1 bulk copy

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // PASS 1 / 2 treat every file as a stream of single records; the pairing happens through the verdict tables
    const bool paired = PASS == 0 && P.mode != 0;
    const bool mmode = PASS == 0 && P.mode == 3;
    const uint32_t lpu = paired ? 8u : 4u;   // lines per unit
    const int nstreams = (paired || PASS == 2) ? 2 : 1;
    const uint32_t nunits2 = PASS == 2 ? ctl->fused_nunits : 0u;   // two files: pairs in this batch (kf2_between)
    RangeCheck rc;
    rc.init(P);
    SwarConsts swar;
    swar.init();

    uint32_t held = 0;   // thread 0: the next ticket (drawn while the previous tile is being staged)
    if (threadIdx.x == 0) held = atomicAdd(&ctl->tile_counter[3], 1u);
    uint32_t mbar_phase = 0;
    if (threadIdx.x < (kFThreads / 32) * 8) (&s_acc[0][0])[threadIdx.x] = 0u;
#if defined(__CUDACC__)
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#endif
    __syncthreads();

    // the tile whose output is staged in s_out and not flushed yet
    bool have_prev = false, p_last = false;
    uint32_t p_tile = 0, p_tot0 = 0, p_tot1 = 0, p_fsel = 0, p_nrec = 0;

#ifdef SK_PHASE_TIMING
    long long t_prev = clock64();
#endif
    while (true) {
        // The ticket was drawn while the previous tile was being staged (S8a), so the bulk copy below goes
        // out without waiting for the atomic's round trip (-2 %).  Not earlier than that: a ticket held a
        // whole tile ahead was measured at 0.38 -> 0.56 ms, because a ticket in the hands of a CTA that
        // is still busy stalls every look-back behind it.
        if (tid == 0) { s_tile = held; s_fail = 0; }
        if (PASS == 3 && tid < 32) s_q[tid] = 0u;
        if (PASS == 4 && tid < 32) s_qacc[tid] = 0u;
        __syncthreads();
        const uint32_t ticket = s_tile;
        // (a tile that gives the batch up poisons the ticket counter -- see kTicketPoison -- so every ticket drawn
        //  after that reads as "no tile left": the launch ends within one tile time instead of working through
        //  10,000 tiles whose output nobody will use.  Tiles are drawn in order, so whatever a tile under way may
        //  be waiting for in a look-back was drawn before the poison and is being finished normally.)
        const bool done = ticket >= num_tiles;
        // two files: tickets are dealt to the files in proportion to their tile counts (ticket t belongs to file 1
        // iff floor((t+1) * tiles_b / num_tiles) > floor(t * tiles_b / num_tiles)), so that both files are walked
        // at the same relative pace; `tile` is the tile's number inside its file
        uint32_t fsel = 0, tile = ticket;
        if (PASS != 0 && !done) {
            const uint32_t lo_ = (uint32_t)((unsigned long long)ticket * tiles_b / num_tiles);
            const uint32_t hi_ = (uint32_t)(((unsigned long long)ticket + 1ull) * tiles_b / num_tiles);
            fsel = hi_ > lo_ ? 1u : 0u;
            tile = fsel ? lo_ : ticket - lo_;
        }
        const DevInput &in = (PASS != 0 && fsel) ? in_b : in_a;
        const uint32_t file_tiles = PASS == 0 ? num_tiles : (fsel ? tiles_b : num_tiles - tiles_b);
        unsigned long long *const status_nl = status_nl_all + (PASS != 0 ? (size_t)fsel * status_stride : 0);
        unsigned long long *const status_out = status_out_all + (PASS != 0 ? (size_t)(2u * fsel) * status_stride : 0);
        unsigned long long *const st_nl[2] = {status_nl, nullptr};
        unsigned long long *const st_out[2] = {status_out, status_out + status_stride};
        const uint32_t nchunks = (in.nbytes + 15u) >> 4;
        const uint32_t t0 = tile * (uint32_t)Cfg::kTile;

        uint32_t mw[(CH + 1) / 2];                               // newline bits, 32 bytes per word
        uint32_t cnt = 0, incl = 0, wbase = 0, n_all = 0, c_t = 0;
        const uint32_t b0 = tid * Cfg::kBytesPerThread;          // region-relative
        bool nl_overflow = false;
        if (!done) {
            // ---- S1: region -> shared memory, one bulk copy (TMA) issued by thread 0 and signalled on an
            // mbarrier; the threads only zero what lies beyond the end of the batch (last tiles)
            const uint32_t have = nchunks - (t0 >> 4);                              // 16-byte chunks from t0 on (>= 1)
            const uint32_t cp_chunks = have < (uint32_t)(Cfg::kRegion / 16) ? have : (uint32_t)(Cfg::kRegion / 16);
#if defined(__CUDACC__)
            if (tid == 0) {
                const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(s_in);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // S8a's reads of s_in come first
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(cp_chunks * 16u + (kSaved ? (uint32_t)kFNlSlot : 0u)) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(in.data + t0), "r"(cp_chunks * 16u), "r"(mb) : "memory");
                if (kSaved)   // the tile's newline positions and line numbers, as the pass before left them
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"((uint32_t)__cvta_generic_to_shared(s_nl)), "l"((fsel ? nlsave_b : nlsave_a) + (size_t)tile * kFNlSlot),
                                   "r"((uint32_t)kFNlSlot), "r"(mb) : "memory");
            }
#else   // host build of the kernels (tests/host_stub/simt): the bulk copy is a memcpy by thread 0
            if (tid == 0) {
                memcpy(s_in, in.data + t0, (size_t)cp_chunks * 16u);
                if (kSaved) memcpy(s_nl, (fsel ? nlsave_b : nlsave_a) + (size_t)tile * kFNlSlot, (size_t)kFNlSlot);
            }
#endif
            for (uint32_t c = cp_chunks + tid; c < (uint32_t)(Cfg::kRegion / 16); c += kFThreads)
                reinterpret_cast<uint4 *>(s_in)[c] = make_uint4(0, 0, 0, 0);
            {   // L2 prefetch, one 128-byte line per thread, of the tile one grid-width ahead: in steady
                // state some CTA (this one, most likely) draws that ticket one tile time from now
                const unsigned long long nb = ((unsigned long long)tile + ((PASS == 0 || PASS == 3) ? gridDim.x : gridDim.x / 2u)) * Cfg::kTile + (unsigned long long)tid * 128u;
#if defined(__CUDACC__)
                if (tid < Cfg::kTile / 128 && nb < in.nbytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(in.data + nb));
#else
                (void)nb;
#endif
            }
#if defined(__CUDACC__)
            {   // wait for the bulk copy (hardware sleep, not a spin), then for the zero fill
                const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
                uint32_t ok = 0;
                while (!ok)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                 : "=r"(ok) : "r"(mb), "r"(mbar_phase) : "memory");
                mbar_phase ^= 1u;
            }
#else
            (void)mbar_phase;
#endif
            __syncthreads();
            SK_TICK(1);   // S1 load

            if (kSaved) {   // found, counted and numbered by the pass before
                const uint4 m = *s_meta;
                c_t = m.y; n_all = m.z;
            }
            // ---- S2: newline masks of this thread's CH*16 contiguous bytes
#pragma unroll
            for (int k = 0; k < (CH + 1) / 2; ++k) mw[k] = 0;
            if (!kSaved) {
                // only the first tile (bytes before the batch start) and the tiles touching the end of
                // the batch have bytes to mask off
                const bool edge = tile == 0 || (unsigned long long)t0 + Cfg::kRegion > in.nbytes;
                const uint32_t lo = (tile == 0) ? in.first : 0u;
                const uint32_t hi = in.nbytes > t0 ? in.nbytes - t0 : 0u;
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const uint4 v = reinterpret_cast<const uint4 *>(s_in)[CH * tid + k];
                    uint32_t mk = newline_mask16(v, swar);
                    if (edge) {
                        const uint32_t cb = b0 + 16u * k;
                        if (cb + 16u > hi) mk &= cb >= hi ? 0u : ((1u << (hi - cb)) - 1u);
                        if (cb < lo) mk &= cb + 16u <= lo ? 0u : (0xffffu << (lo - cb));
                    }
                    mw[k >> 1] |= mk << (16 * (k & 1));
                }
            }
            if (!kSaved) {
#pragma unroll
                for (int k = 0; k < (CH + 1) / 2; ++k) cnt += __popc(mw[k]);
                // ---- S3: ranks
                incl = warp_incl_scan(cnt, lane);
                if (lane == 31) warp_tot[wid] = incl;
                __syncthreads();
                SK_TICK(2);   // S2 masks + warp scan
#pragma unroll
                for (int w = 0; w < kFThreads / 32; ++w) {
                    const uint32_t t = warp_tot[w];
                    if (w < wid) wbase += t;
                    if (w < kFTileThreads / 32) c_t += t;
                    n_all += t;
                }
                nl_overflow = n_all > (uint32_t)kFMaxNl;
                // the tile's newline count goes out before anything else is done with the tile
                const unsigned long long agg[2] = {c_t, 0};
                block_publish(st_nl, tile, agg, 1, epoch, tid);
            }
        }

        SK_TICK(2);   // (+ newline-count publish)
        if (done) {   // no tile left: only the last staged tile remains to be flushed
            if (PASS == 4) {
                if (have_prev && wid >= 4) flush_ordered_tile(outs, s_out, s_fseg, s_flen, s_fdst, P.emu_threads, tid - 128);
            } else if (have_prev && wid >= 4) {
                unsigned long long *const so = status_out_all + (PASS != 0 ? (size_t)(2u * p_fsel) * status_stride : 0);
                unsigned long long *const st_prev[2] = {so, so + status_stride};
                flush_previous_tile(st_prev, p_tile, p_tot0, p_tot1, nstreams, epoch, tid - 128, 4, s_lb, ctl, outs, s_out, p_last,
                                    PASS == 2 ? (int)p_fsel : 0, PASS == 2 ? s_single : nullptr, p_nrec);
            }
            break;
        }

        if (!kSaved && !nl_overflow) {
            // a thread's 16*CH bytes hold about two newlines: one short loop per 32-byte mask word (most
            // words have none, and a warp leaves a word's loop as soon as none of its lanes has one left)
            uint16_t *__restrict__ nl_out = s_nl + (wbase + incl - cnt);
#pragma unroll
            for (int k = 0; k < (CH + 1) / 2; ++k) {
                uint32_t m = mw[k];
                while (m) {
                    *nl_out++ = (uint16_t)(b0 + 32u * k + (uint32_t)__ffs(m) - 1u);
                    m &= m - 1u;
                }
            }
        }
        // ---- S4: global line number of the tile
        uint32_t G;
        if (kSaved) {
            G = s_meta->x;
        } else {
            const unsigned long long agg[2] = {c_t, 0};
            unsigned long long ex[2];
#ifdef SK_KO_LB1
            (void)agg; (void)ex;
            const uint32_t rem = t0 % 325u;   // newlines of a 325-byte record sit at offsets 20, 171, 173, 324
            G = 4u * (t0 / 325u) + (rem > 20u) + (rem > 171u) + (rem > 173u) + (rem > 324u);
#else
            block_walk(st_nl, tile, agg, 1, epoch, tid, s_lb, ex);
            G = (uint32_t)ex[0];
#endif
        }
        if (!kSaved) __syncthreads();   // newline positions visible to every thread (PASS 2, 4: they came with the tile)
        SK_TICK(3);   // S3 positions + S4 look-back #1
        if (PASS == 3 && tq == nullptr && !nl_overflow) {   // (not for the ordered emit) the general path's line index (K1's output), and the line count with the last tile
            for (uint32_t j = (uint32_t)tid; j < c_t; j += kFThreads) {
                if (G + j < in.line_cap) in.line_end[G + j] = t0 + (uint32_t)s_nl[j];
                else ctl->index_overflow = 1u;
            }
            if (tid == 0 && tile == file_tiles - 1u) ctl->nlines[fsel] = G + c_t;
        }
        if ((PASS == 1 || (PASS == 3 && nlsave_a != nullptr)) && !nl_overflow) {   // for PASS 2 / 4: positions (whole 16-byte chunks) and the three numbers
            uint4 *__restrict__ slot = reinterpret_cast<uint4 *>((fsel ? nlsave_b : nlsave_a) + (size_t)tile * kFNlSlot);
            const uint32_t nq = (2u * n_all + 15u) >> 4;
            for (uint32_t c = (uint32_t)tid; c < nq; c += kFThreads) slot[c] = reinterpret_cast<const uint4 *>(s_nl)[c];
            if (tid == 0) slot[kFMaxNl * 2 / 16] = make_uint4(G, c_t, n_all, 0u);
        }

        // ---- S5: units owned by this tile.  Newline j (j < c_t) is global newline G+j; the line
        // after it is line G+j+1; a unit starts at every line that is a multiple of lpu.  Tile 0
        // also owns the unit starting at the batch's first byte ("newline -1").
        int j_s = (int)((lpu - (G + 1u) % lpu) % lpu);
        if (tile == 0) j_s = -1;
        const uint32_t n_units = (j_s < (int)c_t) ? (uint32_t)((int)c_t - 1 - j_s) / lpu + 1u : 0u;
        const uint32_t rpu = paired ? 2u : 1u;
        const uint32_t nrec_t = n_units * rpu;
        bool fail = nl_overflow || nrec_t > (uint32_t)kFThreads / 2u;   // S8a gives every record two lanes
        // S6 gives every record one lane: records live in warps 0..3, warps 4..7 are the flush group
        constexpr int flush_warp0 = 4;

        // ---- deferred S7b/S8b of the PREVIOUS tile, by the flush group, overlapped with S5-S7 of this
        // tile on the record warps: output offsets (look-back #2), then the flush.  s_desc (which aliases
        // s_nl) was consumed by the previous tile's S8a; s_out is only read here and is not written
        // again before the barrier in front of this tile's S8a.  (The flush group's warps never own a
        // record, so they have nothing else to do until that barrier: measured, 22 % of all warp time
        // was spent waiting there.)
        if (PASS == 4) {
            if (have_prev && wid >= flush_warp0) flush_ordered_tile(outs, s_out, s_fseg, s_flen, s_fdst, P.emu_threads, tid - flush_warp0 * 32);
        } else if (have_prev && wid >= flush_warp0) {
            unsigned long long *const so = status_out_all + (PASS != 0 ? (size_t)(2u * p_fsel) * status_stride : 0);
            unsigned long long *const st_prev[2] = {so, so + status_stride};
            flush_previous_tile(st_prev, p_tile, p_tot0, p_tot1, nstreams, epoch, tid - flush_warp0 * 32, kFThreads / 32 - flush_warp0,
                                s_lb, ctl, outs, s_out, p_last, PASS == 2 ? (int)p_fsel : 0, PASS == 2 ? s_single : nullptr, p_nrec);
        }
        have_prev = false;

        // per-record state (lane = record; the mates of a pair sit in adjacent lanes)
        const uint32_t rec = (uint32_t)tid;
        const bool has_rec = !fail && rec < nrec_t;
        bool complete = false;
        uint32_t start = 0, e0 = 0, e1 = 0, e2 = 0, e3 = 0;
        uint32_t recno = 0;                      // two files: the record's number in its file (= its pair's number)
        if (has_rec) {
            const int j = j_s + (int)(lpu * (rec / rpu)) + 4 * (int)(rec % rpu);
            if (PASS != 0) recno = (uint32_t)((int)G + 1 + j) >> 2;
            if ((uint32_t)(j + 4) < n_all) {
                complete = true;
                start = j < 0 ? in.first : (uint32_t)s_nl[j] + 1u;
                e0 = s_nl[j + 1]; e1 = s_nl[j + 2]; e2 = s_nl[j + 3]; e3 = s_nl[j + 4];
            }
        }
        if (paired) {   // a pair is complete only if both mates are (all lanes take part in the shuffle)
            const int mate_complete = __shfl_xor_sync(0xffffffffu, (int)complete, 1);
            complete = complete && mate_complete != 0;
        }
        // an incomplete unit is fine only as the unfinished tail of the batch
        const bool region_to_end = (unsigned long long)t0 + Cfg::kRegion >= in.nbytes;
        if (has_rec && !complete && !region_to_end) fail = true;

        // ---- S6: validate + trim
        TrimOut cut;
        cut.five = -1; cut.three = -1; cut.error = false;
        uint32_t name_len = 0, plus_len = 0, L = 0;
        // PASS 2: a record beyond the shorter file's last one has no mate: it is not part of this batch
        const bool in_batch = PASS != 2 || recno < nunits2;
        unsigned long long mate_verdict = 0;
        if (PASS == 2) {
            if (has_rec && complete && in_batch) {
                name_len = e0 - start;
                L = e1 - e0 - 1u;
                plus_len = e2 - e1 - 1u;
                const unsigned long long v = (fsel ? tab_b : tab_a)[recno];   // PASS 1 validated and trimmed it
                mate_verdict = (fsel ? tab_a : tab_b)[recno];
                if ((v >> 48) & 1u) { cut.five = (int)((v >> 32) & 0xffffu); cut.three = cut.five + (int)((v >> 16) & 0xffffu); }
            }
        } else if (PASS == 4) {
            if (has_rec && complete) {   // validated and trimmed by the index pass
                name_len = e0 - start;
                L = e1 - e0 - 1u;
                plus_len = e2 - e1 - 1u;
                if (recno < tab_cap) {
                    const RecDesc v = desc_out[recno];
                    if (v.route) { cut.five = (int)v.five; cut.three = (int)(v.five + v.nkeep); }
                } else fail = true;
            }
        } else if (has_rec && complete) {
            name_len = e0 - start;
            L = e1 - e0 - 1u;
            plus_len = e2 - e1 - 1u;
            const uint32_t qlen = e3 - e2 - 1u;
            // FQEntry::validate (src/FQEntry.cpp:53-97): any violation is a data error
            if (name_len <= 1u || s_in[start] != '@' || L < 1u || qlen < 1u || qlen != L) fail = true;
            else {
#ifdef SK_KO_S6
                cut.five = 0; cut.three = (int)L;   // nothing trimmed
#else
                cut = lane_sliding_window(s_in, e0 + 1u, L, e2 + 1u, P, rc);
#endif
                if (cut.error) fail = true;
            }
        }
        if (PASS == 1 && has_rec && complete && !fail) {   // the verdict, for PASS 2
            const uint32_t fixed1 = name_len + plus_len + 4u;
            if (recno < tab_cap && fixed1 < 65536u)
                (fsel ? tab_b : tab_a)[recno] = pack_verdict(cut.three >= 0, cut.three >= 0 ? (uint32_t)cut.five : 0u,
                                                              cut.three >= 0 ? (uint32_t)(cut.three - cut.five) : 0u, fixed1);
            else fail = true;
        }
        if (PASS == 3 && has_rec && complete && !fail) {   // the verdict, as k2_trim_only leaves it for k2_trim_route<true>
            if (recno < tab_cap) {
                RecDesc d;
                d.dst_off = 0;
                d.route = cut.three >= 0 ? 1u : 0u;
                d.five = d.route ? (uint32_t)cut.five : 0u;
                d.nkeep = d.route ? (uint32_t)(cut.three - cut.five) : 0u;
                (fsel ? desc_out_b : desc_out)[recno] = d;
                if (tq != nullptr && d.route)
                    atomicAdd(&s_q[(recno + 1u) % (uint32_t)P.emu_threads], name_len + plus_len + 4u + 2u * d.nkeep);
            } else fail = true;
        }
        if (fail) s_fail = 1u;

        // ---- S7: routing + output sizes
        const bool live = has_rec && complete && in_batch;
        const bool keep = live && cut.three >= 0;
        const uint32_t nkeep = keep ? (uint32_t)(cut.three - cut.five) : 0u;
        const uint32_t fixed = name_len + plus_len + 4u;
        uint32_t add0 = 0, add1 = 0;        // bytes for the main stream / the singles stream
        bool nrec_out = false;              // emit as an "N record" (-M)
        int stream = -1;
        // mate's keep flag: the neighbouring lane (interleaved pairs) or the mate's table entry (two files)
        const bool other = PASS == 2 ? ((mate_verdict >> 48) & 1u) != 0
                                     : (paired && __shfl_xor_sync(0xffffffffu, (int)keep, 1) != 0);
        if (kNoEmit) {
            // nothing is emitted in this pass
        } else if (live) {
            if (PASS == 2) {                                        // trim_paired.cpp:543-567, one mate per file
                if (keep && other) { stream = 0; add0 = fixed + 2u * nkeep; }
                else if (keep && P.has_singles) { stream = 1; add1 = fixed + 2u * nkeep; }
                // the mate alone survives: its tile writes it, at this very place of the singles stream
                else if (other && P.has_singles) add1 = (uint32_t)(mate_verdict & 0xffffu) + 2u * (uint32_t)((mate_verdict >> 16) & 0xffffu);
            } else if (!paired) {
                if (keep) { stream = 0; add0 = fixed + 2u * nkeep; }
            } else {
                if (mmode) {                                        // README.md:116-120
                    stream = 0;
                    nrec_out = !keep;
                    add0 = keep ? fixed + 2u * nkeep : fixed + 2u;
                } else if (keep && other) { stream = 0; add0 = fixed + 2u * nkeep; }     // trim_paired.cpp:543-551
                else if (keep && P.has_singles) { stream = 1; add1 = fixed + 2u * nkeep; } // trim_paired.cpp:552-563
            }
        }
        // PASS 4: the records of a tile are staged queue by queue (the reference's dealing: record k of the batch goes to
        // queue (k+1) % N and the queues are written one after the other, src/trim_single.cpp:263,273-274,374-428)
        const uint32_t myq = PASS == 4 ? (recno + 1u) % (uint32_t)P.emu_threads : 0u;
        if (PASS == 4 && rec < 128u) {
            s_rbytes[rec] = add0;
            if (add0) atomicAdd(&s_qacc[myq], add0);
        }
        const uint32_t inc0 = warp_incl_scan(add0, lane);
        const uint32_t inc1 = (paired || PASS == 2) ? warp_incl_scan(add1, lane) : 0u;
        if (lane == 31) { warp_tot2[wid][0] = inc0; warp_tot2[wid][1] = inc1; }
#ifdef SK_PHASE_TIMING
        if (tid == 0) atomicAdd(&g_phase_cycles[0], (unsigned long long)(clock64() - t_prev));   // warp 0's own S5-S7 time
#endif
        __syncthreads();
        SK_TICK(4);   // S5 + S6 + S7 scan (incl. waiting for the other warps)
        uint32_t wb0 = 0, wb1 = 0, tot0 = 0, tot1 = 0;
#pragma unroll
        for (int w = 0; w < kFThreads / 32; ++w) {
            const uint32_t a = warp_tot2[w][0], b = warp_tot2[w][1];
            if (w < wid) { wb0 += a; wb1 += b; }
            tot0 += a; tot1 += b;
        }
        // the staging buffer holds a tile's worth of output; a tile whose (long, untrimmed) records
        // reach far into the halo can exceed it -> general path
        // (PASS 4: every queue's segment starts on a 16-byte boundary of the staging buffer)
        const bool tile_fail = s_fail != 0 || tot0 + tot1 + 64u + (PASS == 4 ? 16u * 32u : 0u) > (uint32_t)Cfg::kOutBytes;
        if (tile_fail) {   // bit 1: the tile simply holds too many records (the host then picks smaller tiles)
            tot0 = 0; tot1 = 0;
            if (tid == 0) {
                atomicOr(&ctl->fast_fail, (nl_overflow || nrec_t > (uint32_t)kFThreads / 2u) ? 3u : 1u);
                atomicOr(&ctl->tile_counter[3], kTicketPoison);
            }
        }
        if (kNoEmit) {   // verdicts are in the table: count the file's records, next tile
            const uint32_t m_live1 = __ballot_sync(0xffffffffu, live);
            if (!tile_fail && lane == 0 && m_live1) s_acc[wid][fsel ? 2 : 0] += (uint32_t)__popc(m_live1);
            if (PASS == 3 && tq != nullptr && !tile_fail && tid < 32) {   // the tile's row, and its share of its group's row
                const uint32_t v = s_q[tid];
                tq[(size_t)tile * 32u + (uint32_t)tid] = v;
                if (v) atomicAdd(&tq[((size_t)num_tiles + 1u + tile / kFTqGroup) * 32u + (uint32_t)tid], v);
            }
            if (tid == 0) held = atomicAdd(&ctl->tile_counter[3], 1u);
            __syncthreads();   // s_fail has been read by everybody before the top of the loop clears it
            continue;
        }
        if (PASS != 4) {   // the tile's output sizes go out now; its own offsets are asked for one tile later
            const unsigned long long agg[2] = {tot0, tot1};
            block_publish(st_out, tile, agg, nstreams, epoch, tid);
        } else {
            // segments of this tile: start in the staging buffer, length, place in the output (the flush group has
            // finished the previous tile: it passed the barrier above)
            if (wid == 0) {
                const uint32_t n = tile_fail ? 0u : s_qacc[lane];
                const uint32_t al = (n + 15u) & ~15u;
                s_fseg[lane] = warp_incl_scan(al, lane) - al;
                s_flen[lane] = n;
                if (lane < P.emu_threads) {   // queue's place + the groups before this tile's + the group's earlier tiles
                    unsigned long long d = (unsigned long long)tq[(size_t)num_tiles * 32u + (uint32_t)lane] +
                                           tq[((size_t)num_tiles + 1u + tile / kFTqGroup) * 32u + (uint32_t)lane];
                    for (uint32_t t2 = tile - tile % kFTqGroup; t2 < tile; ++t2) d += tq[(size_t)t2 * 32u + (uint32_t)lane];
                    s_fdst[lane] = d;
                }
            }
            __syncthreads();
        }
        if (tid == 0) held = atomicAdd(&ctl->tile_counter[3], 1u);   // looked at after S8a
        have_prev = true;
        p_tile = tile; p_tot0 = tot0; p_tot1 = tot1;
        p_fsel = fsel; p_last = tile == file_tiles - 1u; p_nrec = tile_fail ? 0u : nrec_t;
        const uint32_t base1 = (tot0 + 15u) & ~15u;                    // singles staged after the main bytes

        // record descriptors for S8a (two lanes per record); they reuse the newline-position array,
        // which nobody reads after S5
        if (has_rec) {   // every slot S8a may look at is rewritten (x == 0: nothing to emit)
            uint4 dsc;
            uint32_t stage_off = stream == 1 ? wb1 + inc1 - add1 : wb0 + inc0 - add0;
            if (PASS == 4 && stream == 0 && !tile_fail) {   // behind the tile's earlier records of the same queue
                stage_off = s_fseg[myq];
                for (int r2 = (int)rec - P.emu_threads; r2 >= 0; r2 -= P.emu_threads) stage_off += s_rbytes[r2];
            }
            dsc.x = (stream < 0 || tile_fail)
                        ? 0u
                        : (stage_off | (stream == 1 ? 0x80000000u : 0u) |
                           (nrec_out ? 0x40000000u : 0u) | 0x20000000u);
            dsc.y = start | (e0 << 16);
            dsc.z = e1 | (e2 << 16);
            dsc.w = (keep ? (uint32_t)cut.five : 0u) | (nkeep << 16);
            s_desc[rec] = dsc;
            // two files: where this record's own single (if it is one) sits inside the tile's singles range
            if (PASS == 2) s_single[rec] = make_uint2(wb1 + inc1 - add1, (stream == 1 && !tile_fail) ? add1 : 0u);
        }
        __syncthreads();
        SK_TICK(8);   // S7b: totals, publish, descriptors, barrier

        // ---- S8a: two lanes per record copy it into the staging buffer (phase 0; the flush realigns):
        // lane 0 takes [name '\n' seq), lane 1 takes ['\n' line3 '\n' qual '\n')
        {
            const uint32_t r = (uint32_t)tid >> 1, sub = (uint32_t)tid & 1u;
            uint4 dsc = make_uint4(0, 0, 0, 0);
            if (!tile_fail && r < nrec_t) dsc = s_desc[r];   // a failed tile wrote no descriptors
            // Both lanes run the same copies with different arguments (lanes of one warp that took
            // different branches would serialise).  Run A, then -- only in warps where some read is cut
            // at its 5' end -- run B:
            //   lane 0: A = name '\n' [+ seq[0..n) if nothing is cut at 5']      B = seq[five..five+n)
            //   lane 1: A = '\n' line3 '\n' [+ qual[0..n) if nothing is cut at 5']  B = qual[five..five+n)
            // (with five == 0 the kept quality prefix follows line 3 in the input as well); the record's
            // last '\n' is stored separately.
            uint32_t a_dst = 0, a_src = 0, a_len = 0, b_dst = 0, b_src = 0, b_len = 0;
            int nl_at = -1;                       // staging offset of a '\n' to add after the copies
            const bool emit = (dsc.x & 0x20000000u) != 0;
            const bool nrec = (dsc.x & 0x40000000u) != 0;
            const uint32_t r_start = dsc.y & 0xffffu, r_e0 = dsc.y >> 16, r_e1 = dsc.z & 0xffffu, r_e2 = dsc.z >> 16;
            const uint32_t five = dsc.w & 0xffffu, n = dsc.w >> 16;
            const uint32_t nlen = r_e0 - r_start, plen = r_e2 - r_e1 - 1u;
            const uint32_t d = ((dsc.x & 0x80000000u) ? base1 : 0u) + (dsc.x & 0xffffu);
            if (emit) {
                if (sub == 0) {
                    a_dst = d; a_src = r_start;
                    a_len = nlen + 1u + ((five == 0 && !nrec) ? n : 0u);
                    b_dst = d + nlen + 1u; b_src = r_e0 + 1u + five;
                    b_len = (five != 0 && !nrec) ? n : 0u;
                } else if (!nrec) {
                    a_dst = d + nlen + 1u + n; a_src = r_e1;
                    a_len = plen + 2u + (five == 0 ? n : 0u);
                    b_dst = a_dst + plen + 2u; b_src = r_e2 + 1u + five;
                    b_len = five != 0 ? n : 0u;
                    nl_at = (int)(b_dst + n);
                } else {                                             // "N record": name '\n' N '\n' line3 '\n' Qmin '\n'
                    a_dst = d + nlen + 3u; a_src = r_e1 + 1u; a_len = plen + 1u;
                }
            }
            // Experimental (off by default; next round's A/B): the two runs write disjoint bytes, so a lane
            // whose read is cut at the 5' end swaps them -- every lane then moves its LONG piece in the first
            // run and only the short one (name / line 3) in the second, instead of both runs being as long
            // as the longest piece in the warp.
            if (b_len != 0) {
                uint32_t t;
                t = a_dst; a_dst = b_dst; b_dst = t;
                t = a_src; a_src = b_src; b_src = t;
                t = a_len; a_len = b_len; b_len = t;
            }
#ifndef SK_KO_S8A
            smem_copy(s_out, a_dst, s_in, a_src, a_len);
            if (__any_sync(0xffffffffu, b_len != 0)) smem_copy(s_out, b_dst, s_in, b_src, b_len);
#endif
            if (nl_at >= 0) s_out[nl_at] = '\n';
            if (emit && nrec) {
                if (sub == 0) { s_out[d + nlen + 1u] = 'N'; s_out[d + nlen + 2u] = '\n'; }
                else { s_out[a_dst + plen + 1u] = (uint8_t)P.qmin; s_out[a_dst + plen + 2u] = '\n'; }
            }
        }
        SK_TICK(6);   // S8a (thread 0's own copies)
#ifdef SK_PHASE_TIMING
        __syncthreads();   // timing build only: how long the slowest warp takes beyond thread 0
        SK_TICK(9);
#endif

        // ---- bookkeeping: consumed bytes, record count, counters
        if (!tile_fail) {
            // end of the last complete record of this tile (absolute offset in the input buffer)
            uint32_t end = live ? t0 + e3 + 1u : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) end = max(end, __shfl_xor_sync(0xffffffffu, end, o));
            const uint32_t m_live = __ballot_sync(0xffffffffu, live);
            const uint32_t m_keep = __ballot_sync(0xffffffffu, keep);
            uint32_t m_other = 0;
            if (paired) {
                // bit k = keep flag of lane k's mate
                const uint32_t m1 = 0x55555555u;   // lanes of first mates
                m_other = ((m_keep & m1) << 1) | ((m_keep & (m1 << 1)) >> 1);
            }
            if (PASS == 2) m_other = __ballot_sync(0xffffffffu, live && other);
            if (PASS == 2) {
                if (lane == 0 && m_live) {
                    uint32_t *__restrict__ acc = s_acc[wid];
                    acc[fsel ? 3 : 1] = max(acc[fsel ? 3 : 1], end);    // end of the file's last record that is part of a pair
                    if (fsel == 0) {                                        // a pair is counted where its first mate lives
                        acc[4] += (uint32_t)__popc(m_keep & m_other);
                        acc[5] += (uint32_t)__popc(m_live & ~m_keep & ~m_other);
                        acc[6] += (uint32_t)__popc(m_keep & ~m_other);
                        acc[7] += (uint32_t)__popc(m_other & ~m_keep);
                    }
                }
            } else if (lane == 0 && m_live) {   // (only this lane ever touches its warp's row)
                uint32_t *__restrict__ acc = s_acc[wid];
                acc[0] += (uint32_t)__popc(m_live);
                acc[1] = max(acc[1], end);
                if (!paired) {
                    acc[2] += (uint32_t)__popc(m_keep);                 // kept
                    acc[3] += (uint32_t)__popc(m_live & ~m_keep);       // discard
                } else {
                    const uint32_t even = 0x55555555u & m_live;   // one bit per pair (mate 1's lane)
                    const uint32_t k1 = m_keep & even, k2 = m_other & even;
                    acc[4] += (uint32_t)__popc(k1 & k2);                // pairs kept
                    acc[5] += (uint32_t)__popc(even & ~k1 & ~k2);       // pairs discarded
                    acc[6] += (uint32_t)__popc(k1 & ~k2);               // only mate 1 kept
                    acc[7] += (uint32_t)__popc(k2 & ~k1);               // only mate 2 kept
                }
            }
        }
        // the ticket barrier at the top of the loop orders S8a (reads s_in, writes s_out) before the
        // next tile's load (writes s_in) and flush (reads s_out)
    }
    // ---- the CTA's share of the batch summary
    __syncthreads();
    if (tid < 8) {
        uint32_t v = 0;
#pragma unroll
        const bool is_max = tid == 1 || (PASS == 2 && tid == 3);
        for (int w = 0; w < kFThreads / 32; ++w) v = is_max ? max(v, s_acc[w][tid]) : v + s_acc[w][tid];
        if (v && kNoEmit) {            // records of file 0 (row entry 0) and of file 1 (entry 2)
            if (tid == 0) atomicAdd(&ctl->fast_records2[0], v);
            if (tid == 2) atomicAdd(&ctl->fast_records2[1], v);
        } else if (v && PASS == 2 && tid < 4) {
            if (tid == 1) atomicMax(&ctl->fast_consumed2[0], v);
            if (tid == 3) atomicMax(&ctl->fast_consumed2[1], v);
        } else if (v) {
            switch (tid) {
                case 0: atomicAdd(&ctl->fast_records, v); break;
                case 1: atomicMax(&ctl->fast_consumed, v); break;
                case 2: atomicAdd(&ctl->counters[0], (unsigned long long)v); break;          // kept
                case 3: atomicAdd(&ctl->counters[1], (unsigned long long)v); break;          // discard
                case 4: atomicAdd(&ctl->counters[2], 2ull * v); break;                       // kept_p
                case 5: atomicAdd(&ctl->counters[3], 2ull * v); break;                       // discard_p
                case 6: atomicAdd(&ctl->counters[4], (unsigned long long)v);                 // kept_s1
                        atomicAdd(&ctl->counters[7], (unsigned long long)v); break;          // discard_s2
                default: atomicAdd(&ctl->counters[5], (unsigned long long)v);                // kept_s2
                         atomicAdd(&ctl->counters[6], (unsigned long long)v); break;         // discard_s1
            }
        }
    }
}

// Batch summary of the fused path (no line index exists): one thread.
__global__ void kf_finalize(DevInput in, DevParams P, Control *__restrict__ ctl, DevResult *__restrict__ res) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    DevResult r;
    memset(&r, 0, sizeof r);
    for (int s = 0; s < kMaxStreams; ++s) r.out_bytes[s] = ctl->out_bytes[s];
    r.records[0] = ctl->fast_records;
    r.consumed[0] = ctl->fast_consumed;
    for (int k = 0; k < 8; ++k) r.counters[k] = (long long)ctl->counters[k];
    r.index_overflow = (ctl->index_overflow & 2u) | (ctl->fast_fail ? 4u : 0u) | ((ctl->fast_fail & 2u) ? 8u : 0u);
    *res = r;
    Control z;
    memset(&z, 0, sizeof z);
    z.err_key = kNoError;
    *ctl = z;
}

// Between the two passes of a two-file batch: pairs = the smaller of the two record counts; tickets start over
// (a poisoned counter stays poisoned: PASS 2 of a batch PASS 1 gave up ends at once).
__global__ void kf2_between(Control *__restrict__ ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ctl->fused_nunits = min(ctl->fast_records2[0], ctl->fast_records2[1]);
    ctl->tile_counter[3] &= kTicketPoison;
}

// -a N order, between the index pass (kf_fused<CH, 3>) and the ordered emit (kf_fused<CH, 4>).  The index pass left
// tq[t][q] = kept bytes of tile t that go to queue q, and, in the rows behind row `tiles`, the same summed over groups of
// kFTqGroup tiles (atomics; the host zeroes those rows per batch).  One CTA: warp w takes a contiguous run of group rows
// (lane = queue, so a row is one coalesced load), sums it, the warps' sums are scanned through shared memory, and the
// second sweep turns every group row into the exclusive prefix over the groups.  Row `tiles` gets the queues' places in
// the output (the reference writes queue 0, then queue 1, ...: src/trim_single.cpp:374-428); the output size is set and
// the tickets are handed out again (a poisoned counter stays poisoned).  A tile's segment for queue q then starts at
// base[q] + group prefix + the rows of the group's earlier tiles (at most kFTqGroup - 1 loads, in kf_fused<CH, 4>).
__global__ void __launch_bounds__(1024) kfo_offsets(Control *__restrict__ ctl, uint32_t *__restrict__ tq, uint32_t tiles, int nq,
                                                    unsigned long long cap) {
    __shared__ uint32_t part[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t ngroups = (tiles + kFTqGroup - 1u) / kFTqGroup;
    uint32_t *__restrict__ grp = tq + ((size_t)tiles + 1u) * 32u;
    const uint32_t chunk = (ngroups + 31u) / 32u;
    const uint32_t lo = min(ngroups, (uint32_t)w * chunk), hi = min(ngroups, lo + chunk);
    uint32_t sum = 0;
    for (uint32_t g = lo; g < hi; ++g) sum += grp[(size_t)g * 32u + (uint32_t)lane];
    part[w][lane] = sum;
    __syncthreads();
    uint32_t run = 0, total = 0;
    for (int k = 0; k < 32; ++k) {
        const uint32_t v = part[k][lane];
        if (k < w) run += v;
        total += v;
    }
    for (uint32_t g = lo; g < hi; ++g) {
        const uint32_t v = grp[(size_t)g * 32u + (uint32_t)lane];
        grp[(size_t)g * 32u + (uint32_t)lane] = run;
        run += v;
    }
    if (w == 0) {   // the queues' places: exclusive prefix of their totals
        const uint32_t tot = lane < nq ? total : 0u;
        unsigned long long incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        tq[(size_t)tiles * 32u + (uint32_t)lane] = (uint32_t)(incl - tot);
        if (lane == 31) {
            ctl->out_bytes[0] = incl;
            if (incl > cap) {   // reported by the summary as a capacity error; nothing is written
                ctl->index_overflow |= 2u;
                ctl->tile_counter[3] |= kTicketPoison;
            }
            ctl->tile_counter[3] &= kTicketPoison;
        }
    }
}

// Batch summary of a two-file batch.
__global__ void kf2_finalize(Control *__restrict__ ctl, DevResult *__restrict__ res) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    DevResult r;
    memset(&r, 0, sizeof r);
    for (int s = 0; s < kMaxStreams; ++s) r.out_bytes[s] = ctl->out_bytes[s];
    r.records[0] = r.records[1] = ctl->fused_nunits;
    r.consumed[0] = ctl->fast_consumed2[0];
    r.consumed[1] = ctl->fast_consumed2[1];
    for (int k = 0; k < 8; ++k) r.counters[k] = (long long)ctl->counters[k];
    r.index_overflow = (ctl->index_overflow & 2u) | (ctl->fast_fail ? 4u : 0u) | ((ctl->fast_fail & 2u) ? 8u : 0u);
    *res = r;
    Control z;
    memset(&z, 0, sizeof z);
    z.err_key = kNoError;
    *ctl = z;
}

}  // namespace sk
