// kernels_harness.cpp -- TEST INFRASTRUCTURE ONLY (built and run by tests/test_kernels_on_cpu.py).
// Compiles the CUDA kernels of sickle_b200/csrc/*.cuh for the host on top of tests/host_stub/simt/simt_host.h
// (CTAs = OS threads, threads = fibers, warp collectives = exchanges), launches them the way
// sickle_b200/csrc/capi.cu does (launch_fused / launch_general), and compares output streams, counters,
// consumed bytes and the first data error with the CPU oracle's so_run on the same bytes.
//
//   kernels_harness <fastq> <mode 0|2|3 (se, interleaved, -M) or 1 with <fastq2>> <qualtype 1..3> <q> <l> <x> <n>
//                   <has_singles> <path: fused3..fused11|index3..index9|order3..order9|general> <ctas> <first 0..15> [<fastq2>]
// prints one line: "OK ..." / "FASTFAIL ..." (the fused kernel handed the batch to the general path) /
// "MISMATCH ...", exit 0 / 0 / 1.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "simt_host.h"

#include "k1_index.cuh"
#include "k2_trim.cuh"
#include "k3_emit.cuh"
#include "kf_fused.cuh"

#include "sickle_oracle.h"

namespace sk {
thread_local __attribute__((aligned(16))) uint8_t smem[232448];
}

namespace {
std::vector<uint8_t> read_file(const char *path) {
    std::vector<uint8_t> d;
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    return d;
}

template <class T>
T *aligned_zero(size_t n) {
    void *p = nullptr;
    if (posix_memalign(&p, 128, n * sizeof(T) + 256)) abort();
    memset(p, 0, n * sizeof(T) + 256);
    return (T *)p;
}

struct Input {
    uint8_t *buf;      // 16-byte aligned; the batch starts at buf[first]
    sk::DevInput di;
    std::vector<uint8_t> bytes;
};
Input stage_input(const std::vector<uint8_t> &data, uint32_t first, uint32_t line_cap) {
    Input in;
    in.bytes = data;
    // KH_TIGHT=1 (sanitizer runs): exactly what capi.cu guarantees -- 64 readable bytes after the batch, nothing more
    if (getenv("KH_TIGHT")) {
        const size_t n = (data.size() + first + 64 + 15) & ~(size_t)15;
        in.buf = (uint8_t *)aligned_alloc(16, n);
        memset(in.buf, 0, n);
    } else {
        in.buf = aligned_zero<uint8_t>(data.size() + first + 128);
    }
    memset(in.buf, 'x', first);                                     // bytes before the batch: ignored by the kernels
    if (!data.empty()) memcpy(in.buf + first, data.data(), data.size());
    in.di.data = in.buf;
    in.di.first = first;
    in.di.nbytes = (uint32_t)(first + data.size());
    in.di.line_end = aligned_zero<uint32_t>(line_cap + 64);
    in.di.line_cap = line_cap;
    return in;
}

template <int CH>
void run_fused(const sk::DevInput &di, const sk::DevParams &P, sk::Control *ctl, const sk::OutPtrs &op, unsigned ctas, sk::DevResult *res) {
    using Cfg = sk::FusedCfg<CH>;
    const uint32_t tiles = (uint32_t)((di.nbytes + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t cap = tiles + 2;
    unsigned long long *st = aligned_zero<unsigned long long>((size_t)cap * 3 * sk::kWideStatusStride);
    const uint32_t epoch = 5;
    if (tiles) {
        const unsigned grid = std::min<unsigned>(ctas, tiles);
        simt::launch(dim3(grid), dim3(sk::kFThreads), [&] {
            sk::kf_fused<CH>(di, P, ctl, op, st, st + (size_t)cap * sk::kWideStatusStride, cap * sk::kWideStatusStride, tiles, epoch);
        });
    }
    simt::launch(dim3(1), dim3(32), [&] { sk::kf_finalize(di, P, ctl, res); });
    free(st);
}

// two files: both passes of the single-pass kernel, as capi.cu's launch_fused_two_ch does
template <int CH>
void run_fused_two(const sk::DevInput di[2], const sk::DevParams &P, sk::Control *ctl, const sk::OutPtrs &op, unsigned ctas, sk::DevResult *res) {
    using Cfg = sk::FusedCfg<CH>;
    const uint32_t ta = (uint32_t)((di[0].nbytes + Cfg::kTile - 1) / Cfg::kTile), tb = (uint32_t)((di[1].nbytes + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t tiles = ta + tb, cap = std::max(ta, tb) + 2;
    const uint32_t stride = cap * sk::kWideStatusStride;
    unsigned long long *st = aligned_zero<unsigned long long>((size_t)stride * 6);
    const uint32_t tab_cap = (uint32_t)(std::max(di[0].nbytes, di[1].nbytes) / 32 + 64);
    unsigned long long *tab[2] = {aligned_zero<unsigned long long>(tab_cap), aligned_zero<unsigned long long>(tab_cap)};
    uint8_t *nls[2] = {aligned_zero<uint8_t>((size_t)cap * sk::kFNlSlot), aligned_zero<uint8_t>((size_t)cap * sk::kFNlSlot)};   // sized as capi.cu sizes them
    if (tiles) {
        const unsigned grid = std::min<unsigned>(ctas, tiles);
        simt::launch(dim3(grid), dim3(sk::kFThreads), [&] {
            sk::kf_fused<CH, 1>(di[0], P, ctl, op, st, st + 2 * (size_t)stride, stride, tiles, 5u, di[1], tb, tab[0], tab[1], tab_cap, nls[0], nls[1]);
        });
        simt::launch(dim3(1), dim3(32), [&] { sk::kf2_between(ctl); });
        simt::launch(dim3(grid), dim3(sk::kFThreads), [&] {
            sk::kf_fused<CH, 2>(di[0], P, ctl, op, st, st + 2 * (size_t)stride, stride, tiles, 6u, di[1], tb, tab[0], tab[1], tab_cap, nls[0], nls[1]);
        });
    }
    simt::launch(dim3(1), dim3(32), [&] { sk::kf2_finalize(ctl, res); });
    free(st); free(tab[0]); free(tab[1]); free(nls[0]); free(nls[1]);
}

// -a N (N <= 32), single end, as capi.cu's launch_ordered does: index + verdict pass (+ newline positions, bytes per tile
// and queue), kfo_offsets, ordered emit pass, summary
template <int CH>
void run_ordered(const sk::DevInput di[2], const sk::DevParams &P, sk::Control *ctl, const sk::OutPtrs &op, unsigned ctas, sk::DevResult *res) {
    using Cfg = sk::FusedCfg<CH>;
    const uint32_t tiles = (uint32_t)((di[0].nbytes + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t cap = tiles + 2, stride = cap * sk::kWideStatusStride;
    unsigned long long *st = aligned_zero<unsigned long long>((size_t)stride * 3);
    const uint32_t desc_cap = di[0].line_cap / 4 + 1;
    sk::RecDesc *desc = aligned_zero<sk::RecDesc>(desc_cap + 1);
    uint8_t *nls = aligned_zero<uint8_t>((size_t)cap * sk::kFNlSlot);
    uint32_t *tq = aligned_zero<uint32_t>(((size_t)cap + 1 + cap / sk::kFTqGroup + 2) * 32);   // (zeroed: capi.cu memsets the group rows)
    if (tiles) {
        const unsigned grid = std::min<unsigned>(ctas, tiles);
        simt::launch(dim3(grid), dim3(sk::kFThreads), [&] {
            sk::kf_fused<CH, 3>(di[0], P, ctl, op, st, st + (size_t)stride, stride, tiles, 5u, sk::DevInput(), 0u, nullptr, nullptr, desc_cap,
                                nls, nullptr, desc, tq);
        });
        simt::launch(dim3(1), dim3(1024), [&] { sk::kfo_offsets(ctl, tq, tiles, P.emu_threads, op.cap[0]); });
        simt::launch(dim3(grid), dim3(sk::kFThreads), [&] {
            sk::kf_fused<CH, 4>(di[0], P, ctl, op, st, st + (size_t)stride, stride, tiles, 5u, sk::DevInput(), 0u, nullptr, nullptr, desc_cap,
                                nls, nullptr, desc, tq);
        });
    }
    simt::launch(dim3(1), dim3(32), [&] { sk::kf_finalize(di[0], P, ctl, res); });
    free(st); free(desc); free(nls); free(tq);
}

// -a N on one input, as capi.cu's launch_hybrid does: the single-pass kernel's index + verdict pass, then the general
// path's routing, K3 and summary
template <int CH>
void run_hybrid(const sk::DevInput di[2], int n_inputs, const sk::DevParams &P, sk::Control *ctl, const sk::OutPtrs &op, unsigned ctas, sk::DevResult *res) {
    using Cfg = sk::FusedCfg<CH>;
    const uint32_t ta = (uint32_t)((di[0].nbytes + Cfg::kTile - 1) / Cfg::kTile);
    const uint32_t tb = n_inputs == 2 ? (uint32_t)((di[1].nbytes + Cfg::kTile - 1) / Cfg::kTile) : 0u;
    const uint32_t tiles = ta + tb;
    const uint32_t cap = std::max(ta, tb) + 2, stride = cap * sk::kWideStatusStride;
    unsigned long long *st = aligned_zero<unsigned long long>((size_t)stride * 6);
    const uint32_t desc_cap = di[0].line_cap / 4 + 1;
    sk::RecDesc *desc[2] = {aligned_zero<sk::RecDesc>(desc_cap + 1), aligned_zero<sk::RecDesc>(desc_cap + 1)};
    if (tiles) {
        const unsigned grid = std::min<unsigned>(ctas, tiles);
        simt::launch(dim3(grid), dim3(sk::kFThreads), [&] {
            sk::kf_fused<CH, 3>(di[0], P, ctl, op, st, st + 2 * (size_t)stride, stride, tiles, 5u, di[1], tb, nullptr, nullptr, desc_cap,
                                nullptr, nullptr, desc[0], nullptr, desc[1]);
        });
    }
    const uint64_t max_units = ((uint64_t)di[0].nbytes + di[1].nbytes) / 4 + 1;
    const uint64_t t2 = (max_units + sk::kK2UnitsPerTile - 1) / sk::kK2UnitsPerTile;
    const uint32_t k2_cap = (uint32_t)t2 + 2;
    unsigned long long *st2 = aligned_zero<unsigned long long>((size_t)k2_cap * sk::kMaxStreams);
    simt::launch(dim3((unsigned)std::min<uint64_t>(ctas, t2)), dim3(sk::kK2Threads),
                 [&] { sk::k2_trim_route<true>(di[0], di[1], P, ctl, desc[0], desc[1], st2, k2_cap, 5u); });
    simt::launch(dim3(ctas), dim3(sk::kK3Threads), [&] { sk::k3_emit(di[0], di[1], P, ctl, desc[0], desc[1], op); });
    simt::launch(dim3(1), dim3(32), [&] { sk::k_finalize(di[0], di[1], P, ctl, op, res); });
    free(st); free(st2); free(desc[0]); free(desc[1]);
}

void run_general(const sk::DevInput di[2], int n_inputs, const sk::DevParams &P, sk::Control *ctl, const sk::OutPtrs &op, unsigned ctas,
                 sk::DevResult *res) {
    const uint32_t epoch = 9;
    uint32_t line_cap = di[0].line_cap;
    sk::RecDesc *desc[2] = {aligned_zero<sk::RecDesc>(line_cap / 4 + 2), aligned_zero<sk::RecDesc>(line_cap / 4 + 2)};
    for (int i = 0; i < n_inputs; ++i) {
        const uint32_t tiles = (uint32_t)((di[i].nbytes + sk::kK1TileBytes - 1) / sk::kK1TileBytes);
        if (!tiles) continue;
        unsigned long long *st = aligned_zero<unsigned long long>((size_t)(tiles + 2) * sk::kWideStatusStride);
        const sk::DevInput d = di[i];
        simt::launch(dim3(std::min<unsigned>(ctas, tiles)), dim3(sk::kK1Threads), [&] { sk::k1_line_index(d, ctl, i, st, tiles, epoch); });
        free(st);
    }
    const uint64_t max_units = ((uint64_t)di[0].nbytes + di[1].nbytes) / 4 + 1;
    const uint64_t tiles = (max_units + sk::kK2UnitsPerTile - 1) / sk::kK2UnitsPerTile;
    const uint32_t k2_cap = (uint32_t)tiles + 2;
    unsigned long long *st2 = aligned_zero<unsigned long long>((size_t)k2_cap * sk::kMaxStreams);
    simt::launch(dim3((unsigned)std::min<uint64_t>(ctas, tiles)), dim3(sk::kK2Threads),
                 [&] {
                     if (getenv("KH_K2_SPLIT")) return;
                     sk::k2_trim_route<false>(di[0], di[1], P, ctl, desc[0], desc[1], st2, k2_cap, epoch);
                 });
    if (getenv("KH_K2_SPLIT")) {   // the two-kernel form capi.cu uses for batches of long records
        const uint32_t upw = atoi(getenv("KH_K2_SPLIT")) >= 32 ? 32u : sk::kK2aLongUnitsPerTicket;   // as capi.cu: one for long records
        simt::launch(dim3(ctas), dim3(sk::kK2Threads), [&] { sk::k2_trim_only(di[0], di[1], P, ctl, desc[0], desc[1], upw); });
        simt::launch(dim3((unsigned)std::min<uint64_t>(ctas, tiles)), dim3(sk::kK2Threads),
                     [&] { sk::k2_trim_route<true>(di[0], di[1], P, ctl, desc[0], desc[1], st2, k2_cap, epoch); });
    }
    if (getenv("KH_K2_SPLIT") && atoi(getenv("KH_K2_SPLIT")) < 32)   // long records, as capi.cu: K3 with one record per warp
        simt::launch(dim3(ctas), dim3(256), [&] { sk::k3_emit_long(di[0], di[1], P, ctl, desc[0], desc[1], op); });
    else
        simt::launch(dim3(ctas), dim3(sk::kK3Threads), [&] { sk::k3_emit(di[0], di[1], P, ctl, desc[0], desc[1], op); });
    simt::launch(dim3(1), dim3(32), [&] { sk::k_finalize(di[0], di[1], P, ctl, op, res); });
    free(st2); free(desc[0]); free(desc[1]);
}
}  // namespace

int main(int argc, char **argv) {
    if (argc != 12 && argc != 13) { fprintf(stderr, "usage: see the header of kernels_harness.cpp\n"); return 2; }
    const std::vector<uint8_t> d0 = read_file(argv[1]);
    const int mode = atoi(argv[2]), qualtype = atoi(argv[3]);
    sk::DevParams P;
    memset(&P, 0, sizeof P);
    static const int kQ[4][3] = {{0, 4, 60}, {33, 33, 126}, {64, 58, 112}, {64, 64, 110}};   // reference src/sickle.h:85-91
    P.qoff = kQ[qualtype][0]; P.qmin = kQ[qualtype][1]; P.qmax = kQ[qualtype][2];
    P.qthr = atoi(argv[4]); P.lthr = atoi(argv[5]); P.no_fiveprime = atoi(argv[6]); P.trunc_n = atoi(argv[7]);
    P.mode = mode; P.has_singles = atoi(argv[8]);
    P.emu_threads = getenv("KH_THREADS") ? atoi(getenv("KH_THREADS")) : 1;   // reference -a N order inside the batch (general path only)
    const std::string path = argv[9];
    const unsigned ctas = (unsigned)atoi(argv[10]);
    const uint32_t first = (uint32_t)atoi(argv[11]);
    std::vector<uint8_t> d1;
    if (mode == 1) { if (argc != 13) return 2; d1 = read_file(argv[12]); }

    uint32_t line_cap = (uint32_t)((std::max(d0.size(), d1.size()) + 64) / 2 + 64) & ~3u;
    if (getenv("KH_LINE_CAP")) line_cap = (uint32_t)atoi(getenv("KH_LINE_CAP")) & ~3u;        // capacity tests
    Input in0 = stage_input(d0, first, line_cap), in1 = stage_input(d1, mode == 1 ? (first * 7u) % 16u : 0u, line_cap);
    sk::DevInput di[2] = {in0.di, in1.di};
    if (mode != 1) { di[1].data = nullptr; di[1].first = 0; di[1].nbytes = 0; di[1].line_end = nullptr; di[1].line_cap = 0; }
    const size_t cap = d0.size() + d1.size() + 4096;
    const size_t dev_cap = getenv("KH_OUT_CAP") ? (size_t)atoll(getenv("KH_OUT_CAP")) : cap;   // what the kernels are told
    sk::OutPtrs op;
    std::vector<uint8_t *> outs;
    for (int k = 0; k < 3; ++k) {
        if (getenv("KH_TIGHT")) {   // capi.cu allocates cap + 64 bytes per stream, 16-byte aligned
            op.p[k] = (uint8_t *)aligned_alloc(16, (dev_cap + 64 + 15) & ~(size_t)15);
            memset(op.p[k], 0, dev_cap + 64);
        } else {
            op.p[k] = aligned_zero<uint8_t>(cap + 64) + ((k * 5 + first) % 16);   // output buffers at odd phases too
        }
        op.cap[k] = dev_cap;
    }
    if (mode == 0 || mode == 3) { op.cap[1] = op.cap[2] = 0; op.p[1] = op.p[2] = nullptr; }
    if (mode == 2) { op.cap[1] = 0; op.p[1] = nullptr; if (!P.has_singles) { op.cap[2] = 0; op.p[2] = nullptr; } }
    if (mode == 1 && !P.has_singles) { op.cap[2] = 0; op.p[2] = nullptr; }
    sk::Control *ctl = aligned_zero<sk::Control>(1);
    ctl->err_key = sk::kNoError;
    sk::DevResult res;
    memset(&res, 0, sizeof res);

    bool fused = path.rfind("fused", 0) == 0;
    if (fused) {
        const int ch = atoi(path.c_str() + 5);
        if (mode == 1) {
            if (ch == 3) run_fused_two<3>(di, P, ctl, op, ctas, &res);
            else if (ch == 5) run_fused_two<5>(di, P, ctl, op, ctas, &res);
            else if (ch == 7) run_fused_two<7>(di, P, ctl, op, ctas, &res);
            else run_fused_two<9>(di, P, ctl, op, ctas, &res);
        }
        else if (ch == 3) run_fused<3>(di[0], P, ctl, op, ctas, &res);
        else if (ch == 5) run_fused<5>(di[0], P, ctl, op, ctas, &res);
        else if (ch == 7) run_fused<7>(di[0], P, ctl, op, ctas, &res);
        else if (ch == 9) run_fused<9>(di[0], P, ctl, op, ctas, &res);
        else run_fused<11>(di[0], P, ctl, op, ctas, &res);
        if (res.index_overflow & 4u) { printf("FASTFAIL too_many_records=%d\n", (res.index_overflow & 8u) ? 1 : 0); return 0; }
    } else if (path.rfind("order", 0) == 0) {
        const int ch = atoi(path.c_str() + 5);
        if (mode != 0 || P.emu_threads < 2 || P.emu_threads > 32) return 2;
        if (ch == 3) run_ordered<3>(di, P, ctl, op, ctas, &res);
        else if (ch == 5) run_ordered<5>(di, P, ctl, op, ctas, &res);
        else if (ch == 7) run_ordered<7>(di, P, ctl, op, ctas, &res);
        else run_ordered<9>(di, P, ctl, op, ctas, &res);
        if (res.index_overflow & 4u) { printf("FASTFAIL too_many_records=%d\n", (res.index_overflow & 8u) ? 1 : 0); return 0; }
    } else if (path.rfind("index", 0) == 0) {
        const int ch = atoi(path.c_str() + 5);
        const int ni = mode == 1 ? 2 : 1;
        if (ch == 3) run_hybrid<3>(di, ni, P, ctl, op, ctas, &res);
        else if (ch == 5) run_hybrid<5>(di, ni, P, ctl, op, ctas, &res);
        else if (ch == 7) run_hybrid<7>(di, ni, P, ctl, op, ctas, &res);
        else run_hybrid<9>(di, ni, P, ctl, op, ctas, &res);
        if (res.index_overflow & 4u) { printf("FASTFAIL too_many_records=%d\n", (res.index_overflow & 8u) ? 1 : 0); return 0; }
    } else {
        run_general(di, mode == 1 ? 2 : 1, P, ctl, op, ctas, &res);
    }

    if (res.index_overflow & 3u) {   // capi.cu's check_capacity: SK_E_CAPACITY
        printf("OVERFLOW line_index=%d output=%d\n", (int)(res.index_overflow & 1u), (int)((res.index_overflow >> 1) & 1u));
        return 0;
    }
    // ---- the oracle on the same bytes (one batch, input order)
    so_params sp = {qualtype, P.qthr, P.lthr, P.no_fiveprime, P.trunc_n};
    std::vector<char> ob[3];
    char *optr[3];
    size_t ocap[3], olen[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) { ob[k].resize(cap); optr[k] = ob[k].data(); ocap[k] = cap; }
    so_counters ctr;
    so_error err;
    memset(&ctr, 0, sizeof ctr);
    memset(&err, 0, sizeof err);
    // whole units only, as the device sees them: bytes after the last complete unit are left unconsumed
    auto whole = [](const std::vector<uint8_t> &d, size_t lpu, size_t &units) {
        size_t lines = 0;
        for (uint8_t b : d) lines += b == '\n';
        units = lines / lpu;
        return lines;
    };
    size_t u0 = 0, u1 = 0;
    whole(d0, (mode == 2 || mode == 3) ? 8 : 4, u0);
    if (mode == 1) { whole(d1, 4, u1); u0 = std::min(u0, u1); }
    auto cut = [](const std::vector<uint8_t> &d, size_t nlines) {
        size_t p = 0;
        for (size_t k = 0; k < nlines; ++k) p = (size_t)((const uint8_t *)memchr(d.data() + p, '\n', d.size() - p) - d.data()) + 1;
        return p;
    };
    const size_t c0 = cut(d0, u0 * ((mode == 2 || mode == 3) ? 8 : 4)), c1 = mode == 1 ? cut(d1, u0 * 4) : 0;
    int orc = 0;
    if (u0) orc = so_run(mode, &sp, P.emu_threads, (int64_t)1 << 60, P.has_singles, (const char *)d0.data(), c0, (const char *)d1.data(), c1, optr, ocap, olen, &ctr, &err);

    std::string why;
    if (orc != 0 || res.err_kind != 0) {
        if (orc != res.err_kind) why = "error kind " + std::to_string(res.err_kind) + " vs oracle " + std::to_string(orc);
        else if (res.err_record != err.record || res.err_file != err.file) why = "error record";
        else if (orc == SO_ERR_QUAL_RANGE && (res.err_position != err.position || res.err_byte != err.byte)) why = "error position / byte";
        if (why.empty()) { printf("OK error kind=%d record=%lld\n", orc, (long long)err.record); return 0; }
    } else {
        const unsigned long long want_consumed0 = c0 ? c0 + di[0].first : 0, want_consumed1 = c1 ? c1 + di[1].first : 0;
        for (int k = 0; k < 3 && why.empty(); ++k) {
            if (res.out_bytes[k] != olen[k] && op.p[k]) why = "stream " + std::to_string(k) + " size " + std::to_string(res.out_bytes[k]) + " vs " + std::to_string(olen[k]);
            else if (op.p[k] && memcmp(op.p[k], optr[k], olen[k]) != 0) why = "stream " + std::to_string(k) + " bytes";
        }
        const long long want[8] = {ctr.kept, ctr.discard, ctr.kept_p, ctr.discard_p, ctr.kept_s1, ctr.kept_s2, ctr.discard_s1, ctr.discard_s2};
        for (int k = 0; k < 8 && why.empty(); ++k)
            if (res.counters[k] != want[k]) why = "counter " + std::to_string(k);
        if (why.empty() && (res.consumed[0] != want_consumed0 || res.consumed[1] != want_consumed1))
            why = "consumed " + std::to_string(res.consumed[0]) + "/" + std::to_string(res.consumed[1]) + " vs " + std::to_string(want_consumed0) + "/" + std::to_string(want_consumed1);
        if (why.empty() && (res.records[0] != (unsigned long long)ctr.records_in[0] || res.records[1] != (unsigned long long)ctr.records_in[1])) why = "records";
        if (why.empty()) { printf("OK records=%lld out=%zu+%zu+%zu\n", (long long)(ctr.records_in[0] + ctr.records_in[1]), olen[0], olen[1], olen[2]); return 0; }
    }
    printf("MISMATCH %s\n", why.c_str());
    return 1;
}
