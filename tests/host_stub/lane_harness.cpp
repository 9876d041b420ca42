// lane_harness.cpp -- TEST INFRASTRUCTURE ONLY (built and run by tests/test_lane_logic.py).
// Runs the device function sk::lane_sliding_window (sickle_b200/csrc/trim_lane.cuh, compiled for the
// host through tests/host_stub/lane_shim/) on seeded random reads and compares keep / five / three /
// range error with the CPU oracle's so_sliding_window.
//   lane_harness <seed> <reads>     prints "checked <n> kept <k> errors <e> mismatches <m>"
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "trim_lane.cuh"   // the copy next to the shim headers

#include "sickle_oracle.h"


namespace {
struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
    uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};

struct Case {
    std::vector<uint8_t> buf;   // [pad | seq | '\n' '+' '\n' | qual | '\n' | pad]
    uint32_t seq_off, qual_off, L;
    sk::DevParams P;
    so_params sp;
};

const int kQ[4][3] = {{0, 4, 60}, {33, 33, 126}, {64, 58, 112}, {64, 64, 110}};   // reference src/sickle.h:85-91

Case make_case(Rng &r) {
    Case c;
    const int qt = 1 + (int)r.below(3);
    const uint32_t shape = r.below(10);
    uint32_t L = shape < 5 ? 150 : shape < 7 ? 1 + r.below(40) : shape < 9 ? 1 + r.below(600) : 1 + r.below(2500);
    c.L = L;
    const int off = kQ[qt][0], lo = kQ[qt][1] - off, hi = std::min(kQ[qt][2] - off, 45);
    std::vector<uint8_t> seq(L), qual(L);
    const uint32_t pN = r.below(4) == 0 ? 40 : 2000, pn = r.below(6) == 0 ? 60 : 4000;
    for (uint32_t i = 0; i < L; ++i) {
        seq[i] = "ACGT"[r.below(4)];
        if (r.below(pN) == 0) seq[i] = 'N';
        if (r.below(pn) == 0) seq[i] = 'n';
    }
    const uint32_t model = r.below(5);
    const int onset = (int)r.below(L + L / 4 + 1), slope10 = 2 + (int)r.below(30);
    for (uint32_t i = 0; i < L; ++i) {
        int q;
        if (model == 0) q = lo + (int)r.below((uint32_t)(hi - lo + 1));                    // uniform
        else if (model == 1) q = (i / (1 + r.below(3) + L / 7)) % 2 ? 3 + (int)r.below(8) : 30 + (int)r.below(10);   // good / bad stretches
        else {                                                                             // plateau then decay, noisy
            q = 37 - ((int)i > onset ? ((int)i - onset) * slope10 / 10 : 0) + (int)r.below(7) - 3;
            if (i < 3) q = std::min(q, 2 + (int)r.below(40));
        }
        q = std::max(lo, std::min(hi, q));
        qual[i] = (uint8_t)(q + off);
    }
    if (r.below(12) == 0) {   // a byte outside the encoding's range somewhere (a data error if the scan reaches it)
        static const uint8_t odd[] = {0, 10, 13, 32, 127, 128, 200, 255, 57, 63, 111, 113};
        qual[r.below(L)] = odd[r.below(12)];
    }
    static const int qthr[] = {20, 20, 20, 0, 2, 13, 30, 35, 41, 60};
    static const int lthr[] = {20, 20, 0, 1, 5, 50, 151};
    memset(&c.P, 0, sizeof c.P);
    c.P.qoff = off; c.P.qmin = kQ[qt][1]; c.P.qmax = kQ[qt][2];
    c.P.qthr = qthr[r.below(10)];
    c.P.lthr = lthr[r.below(7)];
    c.P.no_fiveprime = r.below(4) == 0;
    c.P.trunc_n = r.below(3) == 0;
    c.P.emu_threads = 1;
    c.sp.qualtype = qt; c.sp.qual_threshold = c.P.qthr; c.sp.length_threshold = c.P.lthr;
    c.sp.no_fiveprime = c.P.no_fiveprime; c.sp.trunc_n = c.P.trunc_n;
    // lay the read out as in a FASTQ buffer, at a random byte phase, followed by another record's worth of bytes
    const uint32_t lead = 8 + r.below(8);
    c.seq_off = lead;
    c.qual_off = lead + L + 3;
    c.buf.assign(c.qual_off + L + 1 + 256, 0);
    for (auto &b : c.buf) b = (uint8_t)(33 + r.below(90));
    memcpy(&c.buf[c.seq_off], seq.data(), L);
    c.buf[c.seq_off + L] = '\n'; c.buf[c.seq_off + L + 1] = '+'; c.buf[c.seq_off + L + 2] = '\n';
    memcpy(&c.buf[c.qual_off], qual.data(), L);
    c.buf[c.qual_off + L] = '\n';
    return c;
}

sk::TrimOut run_lane(const Case &c) {
    sk::RangeCheck rc;
    rc.init(c.P);
    // the buffer must be 4-byte aligned at offset 0 (the function reads aligned words)
    return sk::lane_sliding_window(c.buf.data(), c.seq_off, c.L, c.qual_off, c.P, rc);
}
}  // namespace

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: lane_harness <seed> <reads>\n"); return 2; }
    Rng r{(uint64_t)strtoull(argv[1], nullptr, 10) * 2654435761ULL + 12345};
    const long n = atol(argv[2]);
    long kept = 0, errors = 0, mism = 0, declined = 0;
    for (long i = 0; i < n; ++i) {
        const Case c = make_case(r);
        so_cut cut;
        so_error err;
        memset(&err, 0, sizeof err);
        const int orc = so_sliding_window((const char *)&c.buf[c.seq_off], c.L, (const char *)&c.buf[c.qual_off], c.L, &c.sp, &cut, &err, nullptr);
        const bool want_err = orc == SO_ERR_QUAL_RANGE;
        const bool want_keep = !want_err && cut.three >= 0;
        kept += want_keep; errors += want_err;
        // error = "declined": the function hands the read to the exact warp-wide path.  It must decline when
        // the reference reports a range error, and may decline only if some quality byte of a read it looks
        // at (L >= -l) is outside the encoding's range -- never a clean read.
        bool any_oor = false;
        for (uint32_t j = 0; j < c.L; ++j) any_oor |= c.buf[c.qual_off + j] < c.P.qmin || c.buf[c.qual_off + j] > c.P.qmax;
        const bool may_decline = any_oor && c.L >= (uint32_t)c.P.lthr;
        const sk::TrimOut o = run_lane(c);
        declined += o.error && !want_err;
        const bool keep = !o.error && o.three >= 0;
        const bool ok = o.error ? (want_err || may_decline)
                                : (!want_err && keep == want_keep && (!keep || (o.five == cut.five && o.three == cut.three)));
        if (!ok && ++mism <= 10)
            fprintf(stderr, "MISMATCH read %ld L %u q %d l %d x %d n %d type %d: got err %d five %d three %d, oracle rc %d five %d three %d\n",
                    i, c.L, c.P.qthr, c.P.lthr, c.P.no_fiveprime, c.P.trunc_n, c.sp.qualtype, (int)o.error, o.five, o.three, orc,
                    cut.five, cut.three);
    }
    printf("checked %ld kept %ld errors %ld mismatches %ld declined-without-error %ld\n", n, kept, errors, mism, declined);
    return mism ? 1 : 0;
}
