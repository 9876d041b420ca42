// simt_host.h -- TEST INFRASTRUCTURE ONLY: just enough of the CUDA execution model to run the kernels
// of sickle_b200/csrc/*.cuh on the CPU, so that their integer / byte logic can be checked against the
// oracle by the `-m "not gpu"` tests (tests/test_kernels_on_cpu.py).  Nothing here is used by the product.
//
//   * a CTA is one OS thread; its threads are fibers (ucontext) scheduled round-robin, switching only
//     at barriers, warp collectives and __nanosleep -- so __shared__ maps to thread_local;
//   * __syncthreads / named barriers count arrivals per generation;
//   * __shfl_* / __ballot_sync / __any_sync / __reduce_add_sync are an all-to-all exchange among the
//     lanes named in the mask (every named lane must call, as on the device);
//   * global atomics are the host's; CTAs of a grid run concurrently on their own OS threads, which
//     is what the decoupled look-back (spinning on a predecessor's status word) needs;
//   * PTX in the kernels sits behind `#if defined(__CUDA_ARCH__)` with a plain C++ spelling beside it.
// It makes no attempt at timing, memory spaces or divergence rules: it only has to compute what the
// device computes.
#pragma once

#include <sched.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ thread_local
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct __attribute__((aligned(16))) uint4 { uint32_t x, y, z, w; };
struct dim3 { uint32_t x = 1, y = 1, z = 1; dim3(uint32_t a = 1, uint32_t b = 1, uint32_t c = 1) : x(a), y(b), z(c) {} };
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
typedef void *cudaStream_t;

namespace simt {

struct WarpOp {
    uint32_t mask = 0, arrived = 0, read = 0;
    bool ready = false, used = false;
    uint64_t val[32];
};
struct Warp { WarpOp ops[40]; };   // up to 16 disjoint lane pairs, two generations each
struct Barrier { int count = 0, gen = 0; };
struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<char[]> stack;
    uint3 tid{0, 0, 0};
    bool done = false;
};
struct Cta {
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int cur = 0, alive = 0;
    std::vector<Warp> warps;
    Barrier bars[16];
    uint3 block{0, 0, 0};
    dim3 grid, bdim;
    std::function<void()> body;
};
inline thread_local Cta *cta = nullptr;

inline void yield() {
    Cta *c = cta;
    swapcontext(&c->fibers[(size_t)c->cur].ctx, &c->sched);
}
inline void bar_sync(int id, int nthreads) {
    Barrier &b = cta->bars[id];
    const int g = b.gen;
    if (++b.count >= nthreads) { b.count = 0; b.gen++; return; }
    while (b.gen == g) yield();
}
inline int lane_id() { return (int)(cta->fibers[(size_t)cta->cur].tid.x & 31u); }

// all-to-all among the lanes of `mask`: out[i] = value lane i brought (valid for i in mask)
inline void exchange(uint32_t mask, uint64_t v, uint64_t out[32]) {
    Warp &w = cta->warps[(size_t)cta->cur >> 5];
    const int lane = lane_id();
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "simt: lane %d not in its own mask %08x\n", lane, mask); abort(); }
    WarpOp *op = nullptr;
    for (auto &o : w.ops)
        if (o.used && !o.ready && o.mask == mask && !((o.arrived >> lane) & 1u)) { op = &o; break; }
    if (!op) {
        for (auto &o : w.ops)
            if (!o.used) { op = &o; break; }
        if (!op) { fprintf(stderr, "simt: too many collectives in flight in one warp\n"); abort(); }
        op->used = true; op->ready = false; op->mask = mask; op->arrived = 0; op->read = 0;
    }
    op->val[lane] = v;
    op->arrived |= 1u << lane;
    if (op->arrived == mask) op->ready = true;
    long spins = 0;
    while (!op->ready) {
        yield();
        if (++spins > 50000000L) { fprintf(stderr, "simt: collective with mask %08x never completed (arrived %08x)\n", mask, op->arrived); abort(); }
    }
    memcpy(out, op->val, sizeof op->val);
    op->read |= 1u << lane;
    if (op->read == mask) op->used = false;
}

inline void fiber_entry() {
    Cta *c = cta;
    c->body();
    c->fibers[(size_t)c->cur].done = true;
    c->alive--;
    swapcontext(&c->fibers[(size_t)c->cur].ctx, &c->sched);
}

// Run `body` as a grid: one OS thread per CTA, `block.x` fibers each.
inline void launch(dim3 grid, dim3 block, const std::function<void()> &body, size_t stack_bytes = 256 * 1024) {
    std::vector<std::thread> th;
    for (uint32_t b = 0; b < grid.x; ++b)
        th.emplace_back([=, &body] {
            Cta c;
            c.grid = grid; c.bdim = block; c.block = uint3{b, 0, 0};
            c.body = body;
            c.fibers.resize(block.x);
            c.warps.resize((block.x + 31) / 32);
            c.alive = (int)block.x;
            cta = &c;
            for (uint32_t t = 0; t < block.x; ++t) {
                Fiber &f = c.fibers[t];
                f.tid = uint3{t, 0, 0};
                f.stack.reset(new char[stack_bytes]);
                getcontext(&f.ctx);
                f.ctx.uc_stack.ss_sp = f.stack.get();
                f.ctx.uc_stack.ss_size = stack_bytes;
                f.ctx.uc_link = &c.sched;
                makecontext(&f.ctx, (void (*)())fiber_entry, 0);
            }
            // SIMT_SHUFFLE=<seed>: run the fibers of every scheduling round in a random order instead of by
            // thread index, so that code which only works because a lower-numbered thread ran first (a
            // missing barrier, an unordered shared-memory hand-over) shows up as a mismatch.
            const char *shuf = getenv("SIMT_SHUFFLE");
            uint64_t rng = shuf ? (uint64_t)atoll(shuf) * 0x9E3779B97F4A7C15ULL + b * 0x632BE59BD9B4E019ULL + 1 : 0;
            std::vector<uint32_t> order(block.x);
            for (uint32_t t = 0; t < block.x; ++t) order[t] = t;
            while (c.alive > 0) {
                if (shuf)
                    for (uint32_t i = block.x - 1; i > 0; --i) {
                        rng = rng * 6364136223846793005ULL + 1442695040888963407ULL;
                        std::swap(order[i], order[(uint32_t)((rng >> 33) % (i + 1))]);
                    }
                for (uint32_t k = 0; k < block.x; ++k) {
                    const uint32_t t = order[k];
                    if (c.fibers[t].done) continue;
                    c.cur = (int)t;
                    swapcontext(&c.sched, &c.fibers[t].ctx);
                }
            }
            cta = nullptr;
        });
    for (auto &t : th) t.join();
}

}  // namespace simt

#define threadIdx (simt::cta->fibers[(size_t)simt::cta->cur].tid)
#define blockIdx (simt::cta->block)
#define gridDim (simt::cta->grid)
#define blockDim (simt::cta->bdim)

#define SK_HOST_BAR_SYNC(id, nthreads) simt::bar_sync((id), (nthreads))   // the kernels' named-barrier hook
inline void __syncthreads() { simt::bar_sync(0, (int)simt::cta->bdim.x); }
inline void __syncwarp(uint32_t = 0xffffffffu) {}
inline void __nanosleep(unsigned) { simt::yield(); sched_yield(); }
inline long long clock64() { return 0; }

// ---- warp collectives
template <class T>
inline uint64_t simt_pack(T v) { static_assert(sizeof(T) <= 8, ""); uint64_t u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T>
inline T simt_unpack(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
template <class T>
inline T __shfl_sync(uint32_t mask, T v, int src) {
    uint64_t out[32];
    simt::exchange(mask, simt_pack(v), out);
    return simt_unpack<T>(out[src & 31]);
}
template <class T>
inline T __shfl_xor_sync(uint32_t mask, T v, int x) {
    uint64_t out[32];
    simt::exchange(mask, simt_pack(v), out);
    return simt_unpack<T>(out[(simt::lane_id() ^ x) & 31]);
}
template <class T>
inline T __shfl_up_sync(uint32_t mask, T v, unsigned d) {
    uint64_t out[32];
    simt::exchange(mask, simt_pack(v), out);
    const int lane = simt::lane_id();
    return lane >= (int)d ? simt_unpack<T>(out[lane - (int)d]) : v;
}
inline uint32_t __ballot_sync(uint32_t mask, int pred) {
    uint64_t out[32];
    simt::exchange(mask, (uint64_t)(pred != 0), out);
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i)
        if (((mask >> i) & 1u) && out[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline uint32_t __reduce_min_sync(uint32_t mask, uint32_t v) {
    uint64_t out[32];
    simt::exchange(mask, v, out);
    uint32_t r = 0xffffffffu;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r = std::min(r, (uint32_t)out[i]);
    return r;
}
inline int __all_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline uint32_t __reduce_add_sync(uint32_t mask, uint32_t v) {
    uint64_t out[32];
    simt::exchange(mask, v, out);
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) r += (uint32_t)out[i];
    return r;
}

// ---- integer intrinsics
inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31u)); }
inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)(((((uint64_t)hi << 32) | lo) << (sh & 31u)) >> 32); }
inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
inline int __ffsll(unsigned long long x) { return __builtin_ffsll((long long)x); }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 255u) * ((b >> (8 * k)) & 255u);
    return c;
}
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline T __ldcs(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, T v) { *p = v; }

// ---- atomics (global memory shared by the CTAs' OS threads)
inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline uint32_t atomicOr(uint32_t *p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T>
inline T simt_atomic_minmax(T *p, T v, bool want_max) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while ((want_max ? v > old : v < old) && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline uint32_t atomicMax(uint32_t *p, uint32_t v) { return simt_atomic_minmax(p, v, true); }
inline uint32_t atomicMin(uint32_t *p, uint32_t v) { return simt_atomic_minmax(p, v, false); }
inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { return simt_atomic_minmax(p, v, true); }
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { return simt_atomic_minmax(p, v, false); }

// ---- min / max as CUDA overloads them (mixed signedness converts to unsigned)
using std::max;
using std::min;
inline uint32_t min(uint32_t a, int b) { return a < (uint32_t)b ? a : (uint32_t)b; }
inline uint32_t min(int a, uint32_t b) { return (uint32_t)a < b ? (uint32_t)a : b; }
inline uint32_t max(uint32_t a, int b) { return a > (uint32_t)b ? a : (uint32_t)b; }
inline uint32_t max(int a, uint32_t b) { return (uint32_t)a > b ? (uint32_t)a : b; }
inline unsigned long long min(unsigned long long a, uint32_t b) { return a < b ? a : b; }
inline unsigned long long min(uint32_t a, unsigned long long b) { return a < b ? a : b; }
