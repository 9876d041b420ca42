// copy_harness.cpp -- TEST INFRASTRUCTURE ONLY (built and run by tests/test_copy_logic.py).
// The byte movers of the single-pass kernel (sickle_b200/csrc/sk_copy.cuh, compiled for the host through
// tests/host_stub/lane_shim/) against memcpy: every source / destination phase, lengths 0..700 for the
// staging copy and 0..40000 for the flush, neighbouring bytes must stay untouched (other lanes / other
// tiles own them).   copy_harness <seed> <cases>   prints "smem_copy <n> flush <m> mismatches <k>"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sk_copy.cuh"   // the copy next to the shim headers

namespace {
struct Rng {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(s >> 33); }
    uint32_t below(uint32_t n) { return n ? next() % n : 0; }
};
}  // namespace

int main(int argc, char **argv) {
    if (argc != 3) { fprintf(stderr, "usage: copy_harness <seed> <cases>\n"); return 2; }
    Rng r{(uint64_t)strtoull(argv[1], nullptr, 10) * 0x9E3779B97F4A7C15ULL + 7};
    const long n = atol(argv[2]);
    long mism = 0, n_copy = 0, n_flush = 0;
    alignas(16) static uint8_t in[8192 + 64], out[8192 + 64], want[8192 + 64];
    for (long i = 0; i < n; ++i) {
        for (size_t k = 0; k < sizeof in; ++k) { in[k] = (uint8_t)r.next(); out[k] = (uint8_t)(0xA5 ^ k); }
        memcpy(want, out, sizeof out);
        const uint32_t len = r.below(8) == 0 ? r.below(8) : r.below(700);
        const uint32_t src = r.below(4000), dst = r.below(4000);
        sk::smem_copy(out, dst, in, src, len);
        memcpy(want + dst, in + src, len);
        ++n_copy;
        if (memcmp(out, want, sizeof out) != 0 && ++mism <= 5) fprintf(stderr, "smem_copy MISMATCH dst %u src %u len %u\n", dst, src, len);
    }
    std::vector<uint8_t> stage(65536 + 64), g(70000 + 64), gw(70000 + 64);
    uint8_t *sbase = stage.data() + ((16 - ((uintptr_t)stage.data() & 15)) & 15);   // shared memory is 16-byte aligned
    uint8_t *gbase = g.data() + ((16 - ((uintptr_t)g.data() & 15)) & 15);
    for (long i = 0; i < n / 20 + 1; ++i) {
        for (auto &b : stage) b = (uint8_t)r.next();
        for (size_t k = 0; k < g.size(); ++k) g[k] = (uint8_t)(0x3C ^ k);
        gw = g;
        const uint32_t tot = r.below(6) == 0 ? r.below(40) : r.below(40000);
        const uint32_t sb = 16u * r.below(1000), off = 64 + r.below(20000);
        const int nthreads = r.below(2) ? 96 : 128;
        for (int tid = 0; tid < nthreads; ++tid) sk::flush_realigned(gbase + off, sbase, sb, tot, tid, nthreads);
        memcpy(gw.data() + (gbase - g.data()) + off, sbase + sb, tot);
        ++n_flush;
        if (g != gw && ++mism <= 10) fprintf(stderr, "flush MISMATCH off %u (phase %u) sb %u tot %u threads %d\n", off, off & 15u, sb, tot, nthreads);
    }
    printf("smem_copy %ld flush %ld mismatches %ld\n", n_copy, n_flush, mism);
    return mism ? 1 : 0;
}
