// stub_abi.cpp -- TEST INFRASTRUCTURE ONLY: the C ABI of include/sickle_b200.h answered by the CPU
// oracle (oracle/sickle_oracle.c), so that the host side of the command line (host/*.cpp: option
// parsing, batch cutting, carrying tails, dealing batches to several contexts, ordered output,
// messages, exit codes) can be exercised by the `-m "not gpu"` tests in a container without a GPU.
//
// It is linked only into tests/_build/sickle_hoststub by tests/test_host_logic.py.  The product
// (bin/sickle, sickle_b200/libsickle_b200.so) never sees it: there is no CPU path in the product, and
// nothing measured or shipped runs through this file.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sickle_b200.h"
#include "../../oracle/sickle_oracle.h"

namespace {
thread_local char g_err[256] = "";
void set_err(const char *m) { snprintf(g_err, sizeof g_err, "%s", m); }

struct Slot {
    std::vector<char> in[2], out[3];
    uint64_t r[4] = {0, 0, 0, 0};
    bool busy = false;
};
}  // namespace

struct sk_ctx {
    sk_params p;
    uint64_t slot_bytes;
    int device;
    std::vector<Slot> slots;
};

namespace {
// offset after newline number m (1-based) of buf[0, n); n when there are fewer
uint64_t after_line(const char *buf, uint64_t n, uint64_t m) {
    uint64_t p = 0;
    for (uint64_t k = 0; k < m; ++k) {
        const void *q = memchr(buf + p, '\n', (size_t)(n - p));
        if (!q) return n;
        p = (uint64_t)((const char *)q - buf) + 1;
    }
    return p;
}
uint64_t count_lines(const char *buf, uint64_t n) {
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; ++i) c += buf[i] == '\n';
    return c;
}
}  // namespace

extern "C" {

int sk_abi_version(void) { return SK_ABI_VERSION; }
int sk_device_count(void) {
    const char *e = getenv("SICKLE_STUB_DEVICES");
    return e ? atoi(e) : 1;
}
const char *sk_last_error(void) { return g_err; }

sk_ctx *sk_create(int device, uint64_t slot_bytes, int n_slots, const sk_params *params) {
    if (device < 0 || device >= sk_device_count()) {
        snprintf(g_err, sizeof g_err, "sk_create: CUDA device %d not available (%d devices) -- there is no CPU fallback", device, sk_device_count());
        return nullptr;
    }
    if (!params || n_slots < 1) { set_err("sk_create: bad arguments"); return nullptr; }
    sk_ctx *c = new sk_ctx();
    c->p = *params;
    c->slot_bytes = (slot_bytes + 15) & ~15ull;
    c->device = device;
    c->slots.resize((size_t)n_slots);
    for (auto &s : c->slots) {
        s.in[0].resize(c->slot_bytes + 64);
        if (params->mode == SK_MODE_PE_2FILE) s.in[1].resize(c->slot_bytes + 64);
        for (auto &o : s.out) o.resize(2 * c->slot_bytes + 64);
    }
    return c;
}
void sk_destroy(sk_ctx *ctx) { delete ctx; }

char *sk_in_buffer(sk_ctx *ctx, int slot, int which) {
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size() || which < 0 || which > 1 || ctx->slots[(size_t)slot].in[which].empty()) {
        set_err("sk_in_buffer: bad slot/which");
        return nullptr;
    }
    return ctx->slots[(size_t)slot].in[which].data();
}
uint64_t sk_slot_bytes(const sk_ctx *ctx) { return ctx ? ctx->slot_bytes : 0; }

int sk_upload(sk_ctx *ctx, int slot, int which, uint64_t offset, uint64_t nbytes) {
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size() || which < 0 || which > 1 || offset + nbytes > ctx->slot_bytes) {
        set_err("sk_upload: bad arguments");
        return SK_E_ARG;
    }
    if (ctx->slots[(size_t)slot].busy) { set_err("sk_upload: slot still busy (call sk_wait first)"); return SK_E_ARG; }
    return SK_OK;
}

int sk_submit(sk_ctx *ctx, int slot, uint64_t start0, uint64_t end0, uint64_t start1, uint64_t end1) {
    if (!ctx || slot < 0 || slot >= (int)ctx->slots.size()) { set_err("sk_submit: bad slot"); return SK_E_ARG; }
    Slot &s = ctx->slots[(size_t)slot];
    if (s.busy) { set_err("sk_submit: slot still busy (call sk_wait first)"); return SK_E_ARG; }
    if (start0 > end0 || end0 > ctx->slot_bytes || start1 > end1 || end1 > ctx->slot_bytes) { set_err("sk_submit: bad byte range"); return SK_E_ARG; }
    s.r[0] = start0; s.r[1] = end0; s.r[2] = start1; s.r[3] = end1;
    s.busy = true;
    return SK_OK;
}

int sk_wait(sk_ctx *ctx, int slot, sk_result *res) {
    if (!ctx || !res || slot < 0 || slot >= (int)ctx->slots.size()) { set_err("sk_wait: bad arguments"); return SK_E_ARG; }
    Slot &s = ctx->slots[(size_t)slot];
    if (!s.busy) { set_err("sk_wait: slot has no submitted batch"); return SK_E_ARG; }
    s.busy = false;
    memset(res, 0, sizeof *res);
    const int mode = ctx->p.mode;
    const bool two = mode == SK_MODE_PE_2FILE, inter = mode == SK_MODE_PE_INTER || mode == SK_MODE_PE_INTER_M;
    const char *in[2] = {s.in[0].data() + s.r[0], two ? s.in[1].data() + s.r[2] : nullptr};
    const uint64_t n[2] = {s.r[1] - s.r[0], two ? s.r[3] - s.r[2] : 0};
    const uint64_t lpu = inter ? 8 : 4;
    uint64_t units = count_lines(in[0], n[0]) / lpu;
    if (two) units = std::min<uint64_t>(units, count_lines(in[1], n[1]) / 4);
    res->consumed[0] = after_line(in[0], n[0], units * lpu);
    res->consumed[1] = two ? after_line(in[1], n[1], units * 4) : 0;
    so_params sp = {ctx->p.qualtype, ctx->p.qual_threshold, ctx->p.length_threshold, ctx->p.no_fiveprime, ctx->p.trunc_n};
    char *outp[3] = {s.out[0].data(), s.out[1].data(), s.out[2].data()};
    size_t cap[3] = {s.out[0].size(), s.out[1].size(), s.out[2].size()}, len[3] = {0, 0, 0};
    so_counters ctr;
    so_error err;
    memset(&ctr, 0, sizeof ctr);
    memset(&err, 0, sizeof err);
    if (units)
        so_run(mode, &sp, ctx->p.emulate_threads > 1 ? ctx->p.emulate_threads : 1, (int64_t)1 << 60, ctx->p.has_singles,
               in[0], (size_t)res->consumed[0], in[1], (size_t)res->consumed[1], outp, cap, len, &ctr, &err);
    res->records[0] = (uint64_t)ctr.records_in[0];
    res->records[1] = (uint64_t)ctr.records_in[1];
    res->kept = ctr.kept; res->discard = ctr.discard; res->kept_p = ctr.kept_p; res->discard_p = ctr.discard_p;
    res->kept_s1 = ctr.kept_s1; res->kept_s2 = ctr.kept_s2; res->discard_s1 = ctr.discard_s1; res->discard_s2 = ctr.discard_s2;
    if (err.kind) {
        res->error.kind = err.kind;
        res->error.file = err.file;
        res->error.record = err.record;
        res->error.position = err.position;
        res->error.byte = err.byte;
        const int f = err.file;
        const uint64_t base = f ? s.r[2] : s.r[0];
        uint64_t p = after_line(in[f], n[f], (uint64_t)err.record * 4);
        for (int k = 0; k < 4; ++k) {
            const uint64_t e = after_line(in[f] + p, n[f] - p, 1);
            res->error.line_off[k] = base + p;
            res->error.line_len[k] = e ? e - 1 : 0;
            p += e;
        }
        return SK_OK;
    }
    for (int k = 0; k < 3; ++k) { res->out[k] = s.out[k].data(); res->out_bytes[k] = len[k]; }
    return SK_OK;
}

int sk_trim_device(sk_ctx *, int, const void *, uint64_t, const void *, uint64_t, void *const[3], const uint64_t[3], void *) {
    set_err("host-logic stub: no device path");
    return SK_E_ARG;
}
int sk_result_device(sk_ctx *, int, void *, sk_result *) {
    set_err("host-logic stub: no device path");
    return SK_E_ARG;
}

}  // extern "C"
