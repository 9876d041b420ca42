// trimmer.cpp -- see trimmer.h.  Host logic only: option parsing, file I/O, batch cutting, carrying
// the incomplete tail between batches, ordered output, counters, the reference's messages.
// All per-read work happens on the GPU behind include/sickle_b200.h; there is no CPU trimming here.
#include "trimmer.h"

#include "ref_batcher.h"
#include "unit_cutter.h"

#include <fcntl.h>
#include <getopt.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

namespace host {

bool batch_mode = false;

// Trim_Single::recommended_batch_len / Trim_Paired::recommended_batch_len
// (reference src/trim_single.cpp:194-211, src/trim_paired.cpp:246-263)
static long long recommended_batch_len(unsigned long long file_size, long long b_mib, bool paired) {
    unsigned long long mx = (unsigned)(int)(1024 * 1024 * b_mib);
    if (paired) mx /= 2;
    const unsigned long long rec = file_size / 8;
    if (rec < 20) return 20;
    if (rec > mx) return (long long)mx;
    return (long long)rec;
}

}  // namespace host

using host::ByteSink;
using host::ByteSource;
using host::Totals;

static const char *kTypeNames[4] = {"Phred", "Sanger", "Solexa", "Illumina"};   // reference src/sickle.h:68-73
static const int kQualityConstants[4][3] = {{0, 4, 60}, {33, 33, 126}, {64, 58, 112}, {64, 64, 110}};   // src/sickle.h:85-91

// Print what the reference prints before exit(1): FQEntry::validate (src/FQEntry.cpp:55-94) and
// get_quality_num (src/trim.cpp:130-135).  `record_no` is the 0-based record number in its file.
int Abstract_Trimmer::report_data_error(const sk_result &r, const char *buf0, const char *buf1) {
    const sk_error_info &e = r.error;
    const char *buf = e.file ? buf1 : buf0;
    std::string line[4];
    for (int k = 0; k < 4; ++k) line[k].assign(buf + e.line_off[k], (size_t)e.line_len[k]);
    const long long position = e.record + 1;   // FQEntry::position
    auto err = [](const std::string &s) { fprintf(stderr, "[ERROR] %s\n", s.c_str()); };
    const std::string where = "In " + line[0] + "(line " + std::to_string(position * 4 - 4) + ")";
    switch (e.kind) {
        case SK_DATA_ID_SHORT:
            err(where); err("Sequence ID is to short."); err("ID:" + line[0]); err("Sequence: " + line[1]);
            err("Comment: " + line[2]); err("Qualities: " + line[3]);
            break;
        case SK_DATA_ID_CHAR:
            err(where); err("Invalid char at the beggining of ID."); err("Sequence: " + line[1]);
            err("Comment: " + line[2]); err("Qualities: " + line[3]);
            break;
        case SK_DATA_SEQ_EMPTY: err("Sequence line is empty"); break;
        case SK_DATA_QUAL_EMPTY: err("Quality line is empty."); break;
        case SK_DATA_LEN_MISMATCH:
            err("Sequence and quality lines have different lengths:"); err(line[1]); err(line[3]);
            break;
        case SK_DATA_QUAL_RANGE:
            fprintf(stderr, "ERROR: Quality value (%d) does not fall within correct range for %s encoding.\n", e.byte,
                    kTypeNames[qualtype]);
            fprintf(stderr, "Range for %s encoding: %d-%d\n", kTypeNames[qualtype], kQualityConstants[qualtype][1],
                    kQualityConstants[qualtype][2]);
            fprintf(stderr, "FastQ record: %s\n", line[0].c_str());
            fprintf(stderr, "Quality string: %s\n", line[3].c_str());
            fprintf(stderr, "Quality char: '%c'\n", (char)e.byte);
            fprintf(stderr, "Quality position: %d\n", e.position + 1);
            break;
        default: fprintf(stderr, "[ERROR] unknown data error %d\n", e.kind);
    }
    fflush(stderr);
    return EXIT_FAILURE;
}

namespace {

// `sickle batch` runs many commands in one process: the context (CUDA start-up, pinned slots, device
// buffers) of the previous command is kept and reused when the next one asks for the same thing.
struct CachedCtx {
    sk_ctx *c = nullptr;
    sk_params p{};
    int device = -1, nslots = 0;
    unsigned long long slot = 0;
} g_cache;

// The writer threads read straight out of a context's pinned result buffers (queue_outputs).  Whatever
// path leaves run_device / run_devices -- a data error in batch k while batch k-1 is still being
// written, a failed read, a CUDA error -- the queued writes must be over before those buffers are freed.
void drain_sinks(ByteSink *const *sinks) {
    if (!sinks) return;
    for (int i = 0; i < 3; ++i)
        if (sinks[i] && sinks[i]->is_open()) sinks[i]->drain();
}

// Whether device / pinned buffers are left to process exit instead of being freed one by one (default: yes).
static bool keep_context_at_exit() {
    if (host::batch_mode) return false;   // `sickle batch` runs many commands in one process: nothing may pile up
    const char *e = getenv("SICKLE_B200_KEEP_CONTEXT");
    return e && *e ? atoi(e) != 0 : true;
}

struct Ctx {
    sk_ctx *c = nullptr;
    Totals *tot = nullptr;
    ByteSink *const *sinks = nullptr;   // drained before the context (and its pinned buffers) goes away
    bool reusable = false;   // set once the run finished with every slot idle
    sk_params p{};
    int device = 0, nslots = 0;
    unsigned long long slot = 0;
    bool acquire(int device_, unsigned long long slot_, int nslots_, const sk_params &p_) {
        device = device_; slot = slot_; nslots = nslots_; p = p_;
        if (g_cache.c && g_cache.device == device && g_cache.slot == slot && g_cache.nslots == nslots &&
            memcmp(&g_cache.p, &p, sizeof p) == 0) {
            c = g_cache.c;
            g_cache.c = nullptr;
            return true;
        }
        if (g_cache.c) { sk_destroy(g_cache.c); g_cache.c = nullptr; }
        c = sk_create(device, slot, nslots, &p);
        return c != nullptr;
    }
    ~Ctx() {
        // Buffers and context are left to process exit (the CLI exits right after; unpinning and freeing a 1 GiB slot by
        // hand took 0.2-1.4 s of a 2-4 s run, the exit itself is not slower for it).  SICKLE_B200_KEEP_CONTEXT=0: free them.
        const bool keep = keep_context_at_exit();
        drain_sinks(sinks);
        const double t0 = host::now_s();
        if (c && host::batch_mode && reusable) {
            g_cache.c = c; g_cache.p = p; g_cache.device = device; g_cache.nslots = nslots; g_cache.slot = slot;
        } else if (c && !keep) {
            sk_destroy(c);
        }
        if (tot) tot->t_teardown += host::now_s() - t0;
    }
};

// Output of one batch: queued on the (asynchronous, ordered) sinks straight out of the slot's pinned
// result buffers.  The tickets are waited for before that slot is submitted again.
struct Tickets {
    unsigned long long t[3] = {0, 0, 0};
};
Tickets queue_outputs(const sk_result &r, ByteSink *outs[3]) {
    Tickets k;
    for (int i = 0; i < 3; ++i)
        if (r.out_bytes[i] && outs[i] && outs[i]->is_open()) k.t[i] = outs[i]->write_async(r.out[i], r.out_bytes[i]);
    return k;
}
bool wait_outputs(Tickets &k, ByteSink *outs[3], Totals &tot) {
    const double t0 = host::now_s();
    bool ok = true;
    for (int i = 0; i < 3; ++i)
        if (k.t[i]) { ok = outs[i]->wait(k.t[i]) && ok; k.t[i] = 0; }
    tot.t_write_wait += host::now_s() - t0;
    return ok;
}
void add_totals(Totals &t, const sk_result &r) {
    t.kept += r.kept; t.discard += r.discard;
    t.kept_p += r.kept_p; t.discard_p += r.discard_p;
    t.kept_s1 += r.kept_s1; t.kept_s2 += r.kept_s2;
    t.discard_s1 += r.discard_s1; t.discard_s2 += r.discard_s2;
    t.records[0] += (long long)r.records[0]; t.records[1] += (long long)r.records[1];
    t.kernel_ms += r.kernel_ms;
    t.batches++;
    t.fused_batches += r.fused;
}

unsigned long long env_u64(const char *name, unsigned long long dflt) {
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    return strtoull(e, nullptr, 10);
}

// The reference drops the last character of an unterminated final line (src/GZReader.cpp:81-88):
// overwrite it with the newline the reader would have seen.
inline void patch_eof(char *buf, unsigned long long end) {
    if (end > 0 && buf[end - 1] != '\n') buf[end - 1] = '\n';
}


// Devices to use: SICKLE_B200_DEVICES = comma-separated CUDA device numbers (a number may repeat: two
// contexts on one GPU), or SICKLE_B200_GPUS = N (devices 0..N-1; "all" = every device).  Default: one
// context on device SICKLE_B200_DEVICE (0).
std::vector<int> device_list() {
    std::vector<int> d;
    if (const char *e = getenv("SICKLE_B200_DEVICES")) {
        for (const char *p = e; *p;) {
            char *end = nullptr;
            const long v = strtol(p, &end, 10);
            if (end == p) break;
            d.push_back((int)v);
            p = *end == ',' ? end + 1 : end;
            if (*end && *end != ',') break;
        }
    } else if (const char *g = getenv("SICKLE_B200_GPUS")) {
        const int n = !strcmp(g, "all") ? sk_device_count() : atoi(g);
        for (int i = 0; i < n; ++i) d.push_back(i);
    }
    if (d.empty()) d.push_back((int)env_u64("SICKLE_B200_DEVICE", 0));
    return d;
}

// Several contexts (one per listed device) fed with independent batches.  Batch k runs on context
// k mod G in slot (k / G) mod S.  Every context is driven by its own host thread (the C ABI's rule:
// one thread per context), which submits what is queued for it and waits for its oldest batch; the
// caller fills input buffers, hands batches over with dispatch() and collects results in batch order
// with result().  No data moves between the devices: reads are independent.
class DeviceFarm {
public:
    struct Job {
        long long k = -1;
        uint64_t r[4] = {0, 0, 0, 0};   // start0, end0, start1, end1
    };
    struct Done {
        bool ready = false;
        int rc = SK_OK;
        std::string err;
        sk_result res{};
    };

    ByteSink *const *sinks = nullptr;   // drained before the contexts (and their pinned buffers) go away
    ~DeviceFarm() {
        drain_sinks(sinks);
        for (auto &w : workers_) {
            { std::lock_guard<std::mutex> l(w->mu); w->stop = true; }
            w->cv.notify_all();
        }
        for (auto &w : workers_)
            if (w->th.joinable()) w->th.join();
        const bool keep = keep_context_at_exit();
        if (!keep)
            for (auto &w : workers_)
                if (w->c) sk_destroy(w->c);
    }

    // Contexts are created side by side (pinned allocations dominate: ~0.4 s per GiB each).
    bool create(const std::vector<int> &devices, unsigned long long slot_bytes, int nslots, const sk_params &p, std::string &err) {
        G_ = (int)devices.size();
        S_ = nslots;
        done_.resize((size_t)(G_ * S_));
        std::vector<std::string> errs((size_t)G_);
        std::vector<std::thread> th;
        for (int g = 0; g < G_; ++g) workers_.emplace_back(new Worker());
        for (int g = 0; g < G_; ++g)
            th.emplace_back([&, g] {
                workers_[(size_t)g]->c = sk_create(devices[(size_t)g], slot_bytes, nslots, &p);
                if (!workers_[(size_t)g]->c) errs[(size_t)g] = sk_last_error();
            });
        for (auto &t : th) t.join();
        for (int g = 0; g < G_; ++g)
            if (!workers_[(size_t)g]->c) { err = errs[(size_t)g]; return false; }
        const int n_in = p.mode == SK_MODE_PE_2FILE ? 2 : 1;
        in_.assign((size_t)(G_ * S_ * 2), nullptr);
        for (int g = 0; g < G_; ++g)
            for (int s = 0; s < S_; ++s)
                for (int i = 0; i < n_in; ++i) in_[(size_t)((g * S_ + s) * 2 + i)] = sk_in_buffer(workers_[(size_t)g]->c, s, i);
        for (int g = 0; g < G_; ++g) workers_[(size_t)g]->th = std::thread([this, g] { run(g); });
        return true;
    }

    int depth() const { return G_ * S_; }                       // batches that can be in flight
    int ring(long long k) const { return (int)(k % depth()); }  // index of batch k's slot among all slots
    char *in_buffer(long long k, int which) const { return in_[(size_t)((dev(k) * S_ + slot(k)) * 2 + which)]; }

    void dispatch(long long k, uint64_t s0, uint64_t e0, uint64_t s1, uint64_t e1) {
        { std::lock_guard<std::mutex> l(mu_done_); done_[(size_t)ring(k)] = Done(); }
        Worker &w = *workers_[(size_t)dev(k)];
        Job j;
        j.k = k; j.r[0] = s0; j.r[1] = e0; j.r[2] = s1; j.r[3] = e1;
        { std::lock_guard<std::mutex> l(w.mu); w.jobs.push_back(j); }
        w.cv.notify_all();
    }

    // Result of batch k (valid until its slot is dispatched again); with block == false returns null
    // when the batch is still running.
    const Done *result(long long k, bool block) {
        std::unique_lock<std::mutex> l(mu_done_);
        Done &d = done_[(size_t)ring(k)];
        if (!block && !d.ready) return nullptr;
        cv_done_.wait(l, [&] { return d.ready; });
        return &d;
    }

private:
    struct Worker {
        sk_ctx *c = nullptr;
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::deque<Job> jobs;
        bool stop = false;
    };
    int dev(long long k) const { return (int)(k % G_); }
    int slot(long long k) const { return (int)((k / G_) % S_); }

    void publish(long long k, int rc, const sk_result *res) {
        std::lock_guard<std::mutex> l(mu_done_);
        Done &d = done_[(size_t)ring(k)];
        d.rc = rc;
        if (rc != SK_OK) d.err = sk_last_error();
        if (res) d.res = *res;
        d.ready = true;
        cv_done_.notify_all();
    }

    void run(int g) {
        Worker &w = *workers_[(size_t)g];
        std::deque<Job> inflight;
        while (true) {
            Job j;
            bool submit = false;
            {
                std::unique_lock<std::mutex> l(w.mu);
                w.cv.wait(l, [&] { return !w.jobs.empty() || !inflight.empty() || w.stop; });
                if (!w.jobs.empty() && (int)inflight.size() < S_) { j = w.jobs.front(); w.jobs.pop_front(); submit = true; }
                else if (inflight.empty()) return;   // stop requested, nothing left
            }
            if (submit) {   // keep the device supplied first ...
                const int rc = sk_submit(w.c, slot(j.k), j.r[0], j.r[1], j.r[2], j.r[3]);
                if (rc != SK_OK) publish(j.k, rc, nullptr);
                else inflight.push_back(j);
                continue;
            }
            // ... then wait for the oldest batch of this device
            j = inflight.front();
            inflight.pop_front();
            sk_result r;
            const int rc = sk_wait(w.c, slot(j.k), &r);
            publish(j.k, rc, rc == SK_OK ? &r : nullptr);
        }
    }

    int G_ = 0, S_ = 0;
    std::vector<std::unique_ptr<Worker>> workers_;
    std::vector<char *> in_;
    std::mutex mu_done_;
    std::condition_variable cv_done_;
    std::vector<Done> done_;
};

// Does buf[0, n) hold four complete lines (one whole FASTQ record)?
static bool holds_record(const char *buf, unsigned long long n) {
    const char *p = buf, *end = buf + n;
    for (int k = 0; k < 4; ++k) {
        p = (const char *)memchr(p, '\n', (size_t)(end - p));
        if (!p) return false;
        ++p;
    }
    return true;
}

}  // namespace

int Abstract_Trimmer::run_device(int mode, ByteSource *in0, ByteSource *in1, ByteSink *outs[3], bool has_singles,
                                 Totals &tot) {
    sk_params p;
    memset(&p, 0, sizeof p);
    p.qualtype = qualtype;
    p.qual_threshold = qual_threshold;
    p.length_threshold = length_threshold;
    p.no_fiveprime = no_fiveprime;
    p.trunc_n = trunc_n;
    p.mode = mode;
    p.emulate_threads = threads > 1 ? threads : 1;
    p.has_singles = has_singles ? 1 : 0;
    const std::vector<int> devices = device_list();
    const int device = devices[0];
    const bool two = mode == SK_MODE_PE_2FILE;
    const bool paired = mode != SK_MODE_SE;
    const double t_begin = host::now_s();

    if (sk_device_count() <= 0) {
        fprintf(stderr, "****Error: no usable CUDA device (%s). This build has no CPU path.\n\n", sk_last_error());
        return EXIT_FAILURE;
    }
    if (devices.size() > 1) return run_devices(devices, p, in0, in1, outs, tot);

    // ---------------------------------------------------------------------------------------
    // (A) reference output order requested (-a N, N > 1): batches follow the reference's batch
    //     geometry (src/GZReader.cpp:59-132); one slot.  The output of batch k is written by the
    //     sinks' threads while the host cuts batch k+1 (the cutting is per-line host work).
    // ---------------------------------------------------------------------------------------
    if (p.emulate_threads > 1) {
        const long long bl = host::recommended_batch_len(in0->file_size(), batch_mib, paired);
        // A reference batch counts characters without their newlines (src/GZReader.cpp:68-75): its bytes are at most
        // twice that (one-character lines), so 2 * batch_len always holds a batch of non-empty lines.  Up to 128 MiB
        // that is what the slot gets; above it, the usual FASTQ ratio (records of 64 characters and more: 1/16 extra).
        unsigned long long slot = (unsigned long long)bl <= (128ull << 20) ? 2ull * (unsigned long long)bl + (4ull << 20)
                                                                         : (unsigned long long)bl + (unsigned long long)bl / 16 + (4ull << 20);
        slot = std::min<unsigned long long>(slot, (1ull << 31) - 8192);
        Ctx ctx;
        ctx.tot = &tot;
        ctx.sinks = outs;
        ctx.c = sk_create(device, slot, 1, &p);
        if (!ctx.c) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
        tot.t_init = host::now_s() - t_begin;
        host::RefBatcher b0(in0, bl, (mode == SK_MODE_PE_INTER || mode == SK_MODE_PE_INTER_M) ? 8 : 4);
        host::RefBatcher b1(in1 ? in1 : in0, bl, 4);
        char *h0 = sk_in_buffer(ctx.c, 0, 0), *h1 = two ? sk_in_buffer(ctx.c, 0, 1) : nullptr;
        long long base[2] = {0, 0};
        Tickets pending;   // writes of the previous batch, out of the slot's pinned result buffers
        while (true) {
            const long long n0 = b0.next(h0, slot);
            if (n0 < 0) { fprintf(stderr, "****Error: a reference batch does not fit in a %llu-byte slot.\n\n", slot); return EXIT_FAILURE; }
            if (n0 == 0) break;
            long long n1 = 0;
            if (two) {
                n1 = b1.next(h1, slot);
                if (n1 <= 0) break;
            }
            if (!wait_outputs(pending, outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
            if (sk_submit(ctx.c, 0, 0, (uint64_t)n0, 0, (uint64_t)n1) != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
            sk_result r;
            const double t_w = host::now_s();
            const int wrc = sk_wait(ctx.c, 0, &r);
            tot.t_wait += host::now_s() - t_w;
            if (wrc != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
            if (r.error.kind) { r.error.record += base[r.error.file]; return report_data_error(r, h0, h1); }
            pending = queue_outputs(r, outs);
            add_totals(tot, r);
            base[0] += (long long)r.records[0]; base[1] += (long long)r.records[1];
        }
        if (!wait_outputs(pending, outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
        return EXIT_SUCCESS;
    }

    // Slot size: 64 MiB keeps PCIe and the kernels efficient (a batch is ~1 ms of either) while the
    // pinned buffers stay cheap to allocate (~0.4 s per GiB); small plain files get one small slot's worth.
    unsigned long long slot = env_u64("SICKLE_B200_SLOT_MB", 0) << 20;
    if (!slot) {
        slot = 64ull << 20;
        unsigned long long sz = in0->gzip() ? ~0ull : in0->file_size();
        if (in1) sz = in1->gzip() ? ~0ull : std::max(sz, in1->file_size());
        if (sz < (40ull << 20) && !host::batch_mode) slot = ((sz + sz / 2) | 0xfffffull) + 1;   // >= 4/3 of the file, whole MiB
    }
    const int nslots = two ? 2 : 3;
    Ctx ctx;
    ctx.tot = &tot;
    ctx.sinks = outs;
    if (!ctx.acquire(device, slot, nslots, p)) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
    tot.t_init = host::now_s() - t_begin;
    std::vector<Tickets> tickets((size_t)nslots);
    auto timed_wait = [&](int s_, sk_result *r_) {
        const double t0 = host::now_s();
        const int rc_ = sk_wait(ctx.c, s_, r_);
        tot.t_wait += host::now_s() - t0;
        return rc_;
    };

    // ---------------------------------------------------------------------------------------
    // (B) two input files: pairs are matched by record number.  Each batch is the carried tail of
    //     either file plus new bytes; the next batch is submitted before the previous one's output
    //     is written, so file writes overlap the GPU.
    // ---------------------------------------------------------------------------------------
    if (two) {
        ByteSource *src[2] = {in0, in1};
        unsigned long long carry_len[2] = {0, 0};
        const char *carry_ptr[2] = {nullptr, nullptr};
        bool eof[2] = {false, false};
        long long base[2] = {0, 0};
        int slot_i = 0;
        bool have_prev = false;
        sk_result prev;
        while (true) {
            char *h[2] = {sk_in_buffer(ctx.c, slot_i, 0), sk_in_buffer(ctx.c, slot_i, 1)};
            unsigned long long n[2] = {0, 0};
            bool grew = false;
            for (int i = 0; i < 2; ++i) {
                if (carry_len[i]) memmove(h[i], carry_ptr[i], carry_len[i]);
                n[i] = carry_len[i];
                if (!eof[i]) {
                    const long long r = src[i]->read(h[i] + n[i], slot - n[i]);
                    if (r < 0) { fprintf(stderr, "****Error: read failed\n\n"); return EXIT_FAILURE; }
                    if ((unsigned long long)r < slot - n[i]) eof[i] = true;
                    if (r > 0) grew = true;
                    n[i] += (unsigned long long)r;
                }
                if (eof[i]) patch_eof(h[i], n[i]);
            }
            (void)grew;
            const bool more = n[0] && n[1];   // no pair can be formed once either file is exhausted
            // this slot's previous outputs must have left its pinned result buffers
            if (more && !wait_outputs(tickets[(size_t)slot_i], outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
            if (more && sk_submit(ctx.c, slot_i, 0, n[0], 0, n[1]) != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
            if (have_prev) {   // previous batch's streams are still in the other slot's pinned buffers
                tickets[(size_t)(slot_i ^ 1)] = queue_outputs(prev, outs);
                have_prev = false;
            }
            if (!more) break;
            sk_result r;
            if (timed_wait(slot_i, &r) != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
            if (r.error.kind) { r.error.record += base[r.error.file]; return report_data_error(r, h[0], h[1]); }
            add_totals(tot, r);
            base[0] += (long long)r.records[0]; base[1] += (long long)r.records[1];
            prev = r;
            have_prev = true;
            if (r.consumed[0] == 0 && r.consumed[1] == 0) {
                // a full slot without one whole record is an error; a full slot of one file facing the other
                // file's last, incomplete record is just the end of the pairs
                const bool starved = (n[0] == slot && !holds_record(h[0], n[0])) || (n[1] == slot && !holds_record(h[1], n[1]));
                if (starved) { fprintf(stderr, "****Error: a record does not fit in a %llu-byte slot (raise SICKLE_B200_SLOT_MB).\n\n", slot); return EXIT_FAILURE; }
                break;   // only an incomplete pair is left: dropped, as the reference does at end of file
            }
            for (int i = 0; i < 2; ++i) { carry_ptr[i] = h[i] + r.consumed[i]; carry_len[i] = n[i] - r.consumed[i]; }
            slot_i ^= 1;
        }
        if (have_prev) tickets[(size_t)slot_i] = queue_outputs(prev, outs);   // (slot_i is the slot `prev` ran in)
        for (auto &k : tickets)
            if (!wait_outputs(k, outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
        ctx.reusable = true;
        return EXIT_SUCCESS;
    }

    // ---------------------------------------------------------------------------------------
    // (C) one input stream (se, interleaved pe), pipelined: the bulk of slot k+1 is read at a
    //     headroom offset and uploaded (sk_upload) while batch k runs; once batch k reports how much
    //     it consumed, its unconsumed tail is copied in front of that bulk and batch k+1 is submitted;
    //     only then are batch k's outputs written.
    // ---------------------------------------------------------------------------------------
    const unsigned long long H = std::min<unsigned long long>(slot / 4, env_u64("SICKLE_B200_HEADROOM_MB", 8) << 20);
    const unsigned long long bulk_cap = slot - H;
    struct Pending { int slot; unsigned long long start, end; } pend = {-1, 0, 0};
    int cur = 0;
    bool eof = false;
    long long base = 0;
    while (true) {
        char *h = sk_in_buffer(ctx.c, cur, 0);
        unsigned long long bulk = 0;
        if (!eof) {
            const long long r = in0->read(h + H, bulk_cap);
            if (r < 0) { fprintf(stderr, "****Error: read failed\n\n"); return EXIT_FAILURE; }
            bulk = (unsigned long long)r;
            if (bulk < bulk_cap) eof = true;
            if (eof && bulk) patch_eof(h, H + bulk);   // before the bytes are handed to the copy engine
            if (bulk && sk_upload(ctx.c, cur, 0, H, bulk) != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
        }
        unsigned long long tail = 0;
        sk_result prev;
        bool have_prev = false;
        int prev_slot = -1;
        if (pend.slot >= 0) {
            if (timed_wait(pend.slot, &prev) != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
            char *ph = sk_in_buffer(ctx.c, pend.slot, 0);
            if (prev.error.kind) { prev.error.record += base; return report_data_error(prev, ph, nullptr); }
            add_totals(tot, prev);
            base += (long long)prev.records[0];
            have_prev = true;
            tail = pend.end - pend.start - prev.consumed[0];
            if (tail > H) {
                fprintf(stderr, "****Error: a record (pair) of more than %llu bytes does not fit the carry area (raise SICKLE_B200_HEADROOM_MB).\n\n", H);
                return EXIT_FAILURE;
            }
            if (tail) memcpy(h + H - tail, ph + pend.start + prev.consumed[0], tail);
            const bool stuck = prev.consumed[0] == 0 && bulk == 0;
            prev_slot = pend.slot;
            pend.slot = -1;
            if (stuck) {   // nothing new and nothing consumed: an incomplete record (pair) at end of file
                tickets[(size_t)prev_slot] = queue_outputs(prev, outs);
                break;
            }
        }
        if (bulk + tail > 0) {
            const unsigned long long start = H - tail, end = H + bulk;
            if (eof) patch_eof(h, end);
            // the outputs this slot produced nslots batches ago must have left its pinned result buffers
            if (!wait_outputs(tickets[(size_t)cur], outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
            if (sk_submit(ctx.c, cur, start, end, 0, 0) != SK_OK) { fprintf(stderr, "****Error: %s\n\n", sk_last_error()); return EXIT_FAILURE; }
            pend = {cur, start, end};
            cur = (cur + 1) % nslots;
        }
        // batch k's output is handed to the writer threads only after batch k+1 is on the device
        if (have_prev) tickets[(size_t)prev_slot] = queue_outputs(prev, outs);
        if (pend.slot < 0) break;
    }
    for (auto &k : tickets)
        if (!wait_outputs(k, outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
    ctx.reusable = true;
    return EXIT_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// Several GPUs (SICKLE_B200_DEVICES / SICKLE_B200_GPUS): the input is cut on the host into
// independent batches of whole records (pairs) -- contiguous byte ranges, SURVEY.md 8-e --, batch k
// runs on device k mod G, and the outputs are appended in batch order, so the files are the same
// bytes as with one GPU.  With -a N (N > 1) the batches are the reference's own batches
// (host/ref_batcher.h), which keeps the reference's output order across any number of devices.
// ---------------------------------------------------------------------------------------------
int Abstract_Trimmer::run_devices(const std::vector<int> &devices, const sk_params &p, ByteSource *in0, ByteSource *in1,
                                  ByteSink *outs[3], Totals &tot) {
    const double t_begin = host::now_s();
    const bool two = p.mode == SK_MODE_PE_2FILE;
    const bool inter = p.mode == SK_MODE_PE_INTER || p.mode == SK_MODE_PE_INTER_M;
    const bool ref_order = p.emulate_threads > 1;
    unsigned long long slot;
    long long ref_batch_len = 0;
    if (ref_order) {
        ref_batch_len = host::recommended_batch_len(in0->file_size(), batch_mib, p.mode != SK_MODE_SE);
        slot = (unsigned long long)ref_batch_len <= (128ull << 20) ? 2ull * (unsigned long long)ref_batch_len + (4ull << 20)
                                                                   : (unsigned long long)ref_batch_len + (unsigned long long)ref_batch_len / 16 + (4ull << 20);
        slot = std::min<unsigned long long>(slot, (1ull << 31) - 8192);
    } else {
        slot = env_u64("SICKLE_B200_SLOT_KB", 0) << 10;   // (small slots: tests)
        if (!slot) slot = env_u64("SICKLE_B200_SLOT_MB", 64) << 20;
        slot = std::min<unsigned long long>(std::max<unsigned long long>(slot, 4096), (1ull << 31) - 8192);
    }
    const int nslots = ref_order && slot > (256ull << 20) ? 1 : 2;
    DeviceFarm farm;
    farm.sinks = outs;
    {
        std::string err;
        if (!farm.create(devices, slot, nslots, p, err)) { fprintf(stderr, "****Error: %s\n\n", err.c_str()); return EXIT_FAILURE; }
    }
    tot.t_init = host::now_s() - t_begin;

    const int depth = farm.depth();
    std::vector<Tickets> tickets((size_t)depth);
    std::vector<uint64_t> cut0((size_t)depth, 0), cut1((size_t)depth, 0);   // submitted sizes, checked against `consumed`
    long long dispatched = 0, retired = 0;
    long long base[2] = {0, 0};

    // Take the result of the oldest batch not yet collected: errors first, then its streams go to the
    // writers (in batch order).  0 = done, 1 = nothing ready (only when !block), -1 = failed (exit code in rc).
    int rc = EXIT_SUCCESS;
    auto retire = [&](bool block) -> int {
        const double t0 = host::now_s();
        const DeviceFarm::Done *d = farm.result(retired, block);
        if (block) tot.t_wait += host::now_s() - t0;
        if (!d) return 1;
        const int ring = farm.ring(retired);
        if (d->rc != SK_OK) { fprintf(stderr, "****Error: %s\n\n", d->err.c_str()); rc = EXIT_FAILURE; return -1; }
        sk_result r = d->res;
        if (r.error.kind) {
            r.error.record += base[r.error.file];
            rc = report_data_error(r, farm.in_buffer(retired, 0), two ? farm.in_buffer(retired, 1) : nullptr);
            return -1;
        }
        if (r.consumed[0] != cut0[(size_t)ring] || (two && r.consumed[1] != cut1[(size_t)ring])) {
            fprintf(stderr, "****Error: internal: batch %lld was cut at %llu / %llu bytes but the device consumed %llu / %llu.\n\n", retired,
                    (unsigned long long)cut0[(size_t)ring], (unsigned long long)cut1[(size_t)ring],
                    (unsigned long long)r.consumed[0], (unsigned long long)r.consumed[1]);
            rc = EXIT_FAILURE;
            return -1;
        }
        tickets[(size_t)ring] = queue_outputs(r, outs);
        add_totals(tot, r);
        base[0] += (long long)r.records[0]; base[1] += (long long)r.records[1];
        ++retired;
        return 0;
    };
    // Make batch k's slot usable: the batch that ran in it `depth` batches ago has been collected and
    // its output has left the slot's pinned result buffers.
    auto claim = [&](long long k) -> bool {
        while (retired + depth <= k)
            if (retire(true) != 0) return false;
        if (!wait_outputs(tickets[(size_t)farm.ring(k)], outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); rc = EXIT_FAILURE; return false; }
        return true;
    };
    auto dispatch = [&](long long k, uint64_t n0, uint64_t n1) {
        cut0[(size_t)farm.ring(k)] = n0; cut1[(size_t)farm.ring(k)] = n1;
        farm.dispatch(k, 0, n0, 0, n1);
        dispatched = k + 1;
    };
    auto collect_ready = [&]() -> bool {
        while (retired < dispatched) {
            const int s = retire(false);
            if (s < 0) return false;
            if (s > 0) break;
        }
        return true;
    };

    if (ref_order) {
        host::RefBatcher b0(in0, ref_batch_len, inter ? 8 : 4);
        host::RefBatcher b1(in1 ? in1 : in0, ref_batch_len, 4);
        for (long long k = 0;; ++k) {
            if (!claim(k)) return rc;
            const long long n0 = b0.next(farm.in_buffer(k, 0), slot);
            if (n0 < 0) { fprintf(stderr, "****Error: a reference batch does not fit in a %llu-byte slot.\n\n", slot); return EXIT_FAILURE; }
            if (n0 == 0) break;
            long long n1 = 0;
            if (two) {
                n1 = b1.next(farm.in_buffer(k, 1), slot);
                if (n1 <= 0) break;
            }
            dispatch(k, (uint64_t)n0, (uint64_t)n1);
            if (!collect_ready()) return rc;
        }
    } else {
        host::UnitStream s0(in0, inter ? 8 : 4), s1(in1 ? in1 : in0, 4);
        for (long long k = 0;; ++k) {
            if (!claim(k)) return rc;
            if (!s0.fill(farm.in_buffer(k, 0), slot) || (two && !s1.fill(farm.in_buffer(k, 1), slot))) {
                fprintf(stderr, "****Error: read failed\n\n");
                return EXIT_FAILURE;
            }
            const unsigned long long units = two ? std::min(s0.units(), s1.units()) : s0.units();
            if (units == 0) {
                // nothing whole is left (what remains is dropped, as the reference does at end of file),
                // unless a full buffer holds no complete record (pair)
                const bool starved = (s0.units() == 0 && s0.full() && !s0.eof()) || (two && s1.units() == 0 && s1.full() && !s1.eof());
                if (starved) { fprintf(stderr, "****Error: a record does not fit in a %llu-byte slot (raise SICKLE_B200_SLOT_MB).\n\n", slot); return EXIT_FAILURE; }
                break;
            }
            const unsigned long long n0 = s0.cut(units), n1 = two ? s1.cut(units) : 0;
            dispatch(k, n0, n1);
            if (!collect_ready()) return rc;
        }
    }
    while (retired < dispatched)
        if (retire(true) != 0) return rc;
    for (auto &k : tickets)
        if (!wait_outputs(k, outs, tot)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
    return EXIT_SUCCESS;
}


// =============================================================================================
// sickle se      (reference src/trim_single.cpp)
// =============================================================================================
static struct option single_long_options[] = {
    {"fastq-file", required_argument, 0, 'f'}, {"output-file", required_argument, 0, 'o'},
    {"qual-type", required_argument, 0, 't'},  {"qual-threshold", required_argument, 0, 'q'},
    {"length-threshold", required_argument, 0, 'l'}, {"no-fiveprime", no_argument, 0, 'x'},
    {"discard-n", no_argument, 0, 'n'},        {"gzip-output", no_argument, 0, 'g'},
    {"quiet", no_argument, 0, 'z'},            {"threads", no_argument, 0, 'a'},
    {"batch", no_argument, 0, 'b'},            {"help", no_argument, 0, CHAR_MIN - 2},
    {"version", no_argument, 0, CHAR_MIN - 3}, {NULL, 0, NULL, 0}};

static void print_version_and_exit() {
    fprintf(stdout, "%s version %0.3f\nCopyright (c) 2011 The Regents of University of California, Davis Campus.\n"
                    "%s is free software and comes with ABSOLUTELY NO WARRANTY.\nDistributed under the MIT License.\n\n"
                    "Written by %s\n", PROGRAM_NAME, (double)SICKLE_VERSION, PROGRAM_NAME,
            "Nikhil Joshi, UC Davis Bioinformatics Core\n");
    exit(EXIT_SUCCESS);
}

static int parse_qualtype(const char *s) {
    if (!strcmp(s, "illumina")) return SK_QUAL_ILLUMINA;
    if (!strcmp(s, "solexa")) return SK_QUAL_SOLEXA;
    if (!strcmp(s, "sanger")) return SK_QUAL_SANGER;
    return -1;
}

void Trim_Single::usage(int status, char const *msg) {
    fprintf(stderr, "\nUsage: %s se [options] -f <fastq sequence file> -t <quality type> -o <trimmed fastq file>\n\
\n\
Options:\n\
-f, --fastq-file, Input fastq file (required)\n\
-t, --qual-type, Type of quality values (solexa (CASAVA < 1.3), illumina (CASAVA 1.3 to 1.7), sanger (which is CASAVA >= 1.8)) (required)\n\
-o, --output-file, Output trimmed fastq file (required)\n", PROGRAM_NAME);
    fprintf(stderr, "-q, --qual-threshold, Threshold for trimming based on average quality in a window. Default 20.\n\
-l, --length-threshold, Threshold to keep a read based on length after trimming. Default 20.\n\
-x, --no-fiveprime, Don't do five prime trimming.\n\
-n, --trunc-n, Truncate sequences at position of first N.\n\
-g, --gzip-output, Output gzipped files.\n\
-a, --threads, Number of threads to use. Default and minimum: Available cores - 1.\n\
-b, --batch, maximum MB of data to read from the input file at each cycle.\n\
\tThe greater the value, the greater the memory usage can be. The value, multiplied by 1024^2, must be \n\
\tbigger than the lenght of the longest read. Minimum 1. Default: 512.\n\
--quiet, Don't print out any trimming information\n\
--help, display this help and exit\n\
--version, output version information and exit\n\n");
    if (msg) fprintf(stderr, "%s\n\n", msg);
    exit(status);
}

int Trim_Single::parse_args(int argc, char *argv[]) {
    int optc;
    while (true) {
        int option_index = 0;
        optc = getopt_long(argc, argv, "df:t:o:q:a:b:l:zxng", single_long_options, &option_index);
        if (optc == -1) break;
        switch (optc) {
            case 'f': infn = strdup(optarg); break;
            case 't':
                qualtype = parse_qualtype(optarg);
                if (qualtype < 0) { fprintf(stderr, "Error: Quality type '%s' is not a valid type.\n", optarg); return EXIT_FAILURE; }
                break;
            case 'o': outfn = strdup(optarg); break;
            case 'q':
                qual_threshold = atoi(optarg);
                if (qual_threshold < 0) { fprintf(stderr, "Quality threshold must be >= 0\n"); return EXIT_FAILURE; }
                break;
            case 'l':
                length_threshold = atoi(optarg);
                if (length_threshold < 0) { fprintf(stderr, "Length threshold must be >= 0\n"); return EXIT_FAILURE; }
                break;
            case 'x': no_fiveprime = 1; break;
            case 'n': trunc_n = 1; break;
            case 'g': gzip_output = 1; break;
            case 'z': quiet = 1; break;
            case 'd': debug = 1; break;
            case 'a': threads = atoi(optarg); threads_given = true; break;
            case 'b': batch_mib = atoi(optarg); break;
            case CHAR_MIN - 2: usage(EXIT_SUCCESS, NULL); break;
            case CHAR_MIN - 3: print_version_and_exit(); break;
            default: usage(EXIT_FAILURE, NULL); break;
        }
    }
    if (qualtype == -1 || !infn || !outfn) usage(EXIT_FAILURE, "****Error: Must have quality type, input file, and output file.");
    if (!strcmp(infn, outfn)) { fprintf(stderr, "****Error: Input file is same as output file.\n\n"); return EXIT_FAILURE; }
    return 0;
}

int Trim_Single::trim_main() {
    const double t_start = host::now_s();
    ByteSource in;
    if (!in.open(infn)) { fprintf(stderr, "****Error: Could not open input file '%s'.\n\n", infn); return EXIT_FAILURE; }
    ByteSink out;
    if (!out.open(outfn, gzip_output != 0)) { fprintf(stderr, "****Error: Could not open output file '%s'.\n\n", outfn); return EXIT_FAILURE; }
    ByteSink *outs[3] = {&out, nullptr, nullptr};
    Totals t;
    const int rc = run_device(SK_MODE_SE, &in, nullptr, outs, false, t);
    const bool closed = out.close();
    if (rc != EXIT_SUCCESS) return rc;
    if (!closed) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
    if (!quiet)
        fprintf(stdout, "\nSE input file: %s\n\nTotal FastQ records: %lld\nFastQ records kept: %lld\nFastQ records discarded: %lld\n\n",
                infn, t.kept + t.discard, t.kept, t.discard);
    if (debug)
        fprintf(stderr, "[sickle_b200] batches %lld (fused %lld), kernel %.3f ms | host: init %.3f s, read %.3f s%s, device wait %.3f s, "
                        "write wait %.3f s (writer busy %.3f s), teardown %.3f s, total %.3f s\n",
                t.batches, t.fused_batches, t.kernel_ms, t.t_init, in.read_seconds(), in.gzip() ? " (gzip)" : "", t.t_wait,
                t.t_write_wait, out.busy_seconds(), t.t_teardown, host::now_s() - t_start);
    return EXIT_SUCCESS;
}

// =============================================================================================
// sickle pe      (reference src/trim_paired.cpp)
// =============================================================================================
static struct option paired_long_options[] = {
    {"qual-type", required_argument, 0, 't'},     {"pe-file1", required_argument, 0, 'f'},
    {"pe-file2", required_argument, 0, 'r'},      {"pe-interleaved", required_argument, 0, 'c'},
    {"output-pe1", required_argument, 0, 'o'},    {"output-pe2", required_argument, 0, 'p'},
    {"output-single", required_argument, 0, 's'}, {"output-interleaved", required_argument, 0, 'm'},
    {"output-combo-all", required_argument, 0, 'M'}, {"qual-threshold", required_argument, 0, 'q'},
    {"length-threshold", required_argument, 0, 'l'}, {"no-fiveprime", no_argument, 0, 'x'},
    {"truncate-n", no_argument, 0, 'n'},          {"gzip-output", no_argument, 0, 'g'},
    {"quiet", no_argument, 0, 'z'},               {"threads", no_argument, 0, 'a'},
    {"batch", no_argument, 0, 'b'},               {"help", no_argument, 0, CHAR_MIN - 2},
    {"version", no_argument, 0, CHAR_MIN - 3},    {NULL, 0, NULL, 0}};

void Trim_Paired::usage(int status, char const *msg) {
    fprintf(stderr, "\nIf you have separate files for forward and reverse reads:\n");
    fprintf(stderr, "Usage: %s pe [options] -f <paired-end forward fastq file> -r <paired-end reverse fastq file> -t <quality type> -o <trimmed PE forward file> -p <trimmed PE reverse file> -s <trimmed singles file>\n\n", PROGRAM_NAME);
    fprintf(stderr, "If you have one file with interleaved forward and reverse reads:\n");
    fprintf(stderr, "Usage: %s pe [options] -c <interleaved input file> -t <quality type> -m <interleaved trimmed paired-end output> -s <trimmed singles file>\n\n\
If you have one file with interleaved reads as input and you want ONLY one interleaved file as output:\n\
Usage: %s pe [options] -c <interleaved input file> -t <quality type> -M <interleaved trimmed output>\n\n", PROGRAM_NAME, PROGRAM_NAME);
    fprintf(stderr, "Options:\n\
Paired-end separated reads\n\
--------------------------\n\
-f, --pe-file1, Input paired-end forward fastq file (Input files must have same number of records)\n\
-r, --pe-file2, Input paired-end reverse fastq file\n\
-o, --output-pe1, Output trimmed forward fastq file\n\
-p, --output-pe2, Output trimmed reverse fastq file. Must use -s option.\n\n\
Paired-end interleaved reads\n\
----------------------------\n");
    fprintf(stderr, "-c, --pe-interleaved, Combined (interleaved) input paired-end fastq\n\
-m, --output-interleaved, Output combined (interleaved) paired-end fastq file. Must use -s option.\n\
-M, --output-combo-all, Output combined (interleaved) paired-end fastq file with any discarded read written to output file as a single N. Cannot be used with the -s option.\n\
--------------\n\
-t, --qual-type, Type of quality values (solexa (CASAVA < 1.3), illumina (CASAVA 1.3 to 1.7), sanger (which is CASAVA >= 1.8)) (required)\n");
    fprintf(stderr, "-s, --output-single, Output trimmed singles fastq file\n\
-q, --qual-threshold, Threshold for trimming based on average quality in a window. Default 20.\n\
-l, --length-threshold, Threshold to keep a read based on length after trimming. Default 20.\n\
-x, --no-fiveprime, Don't do five prime trimming.\n\
-n, --truncate-n, Truncate sequences at position of first N.\n\
-a, --threads, Number of threads to use. Default and minimum: Available cores - 1.\n\
-b, --batch, maximum MB of data to read from the input file at each cycle.\n\
\tThe greater the value, the greater the memory usage can be. The value, multiplied by 1024^2, must be \n\
\tbigger than the lenght of the longest read. Minimum 1. Default: 512.\n");
    fprintf(stderr, "-g, --gzip-output, Output gzipped files.\n--quiet, do not output trimming info\n\
--help, display this help and exit\n\
--version, output version information and exit\n\n");
    if (msg) fprintf(stderr, "%s\n\n", msg);
    exit(status);
}

int Trim_Paired::parse_args(int argc, char *argv[]) {
    int optc;
    while (true) {
        int option_index = 0;
        optc = getopt_long(argc, argv, "df:r:c:t:o:p:m:M:s:q:a:b:l:xng", paired_long_options, &option_index);
        if (optc == -1) break;
        switch (optc) {
            case 'f': infn = strdup(optarg); break;
            case 'r': infn2 = strdup(optarg); break;
            case 'c': infnc = strdup(optarg); break;
            case 't':
                qualtype = parse_qualtype(optarg);
                if (qualtype < 0) { fprintf(stderr, "Error: Quality type '%s' is not a valid type.\n", optarg); return EXIT_FAILURE; }
                break;
            case 'o': outfn = strdup(optarg); break;
            case 'p': outfn2 = strdup(optarg); break;
            case 'm': outfnc = strdup(optarg); break;
            case 'M': outfnM = strdup(optarg); break;   // absent from the fork's switch (usage + exit 1 there); README.md:116-120
            case 's': sfn = strdup(optarg); break;
            case 'q':
                qual_threshold = atoi(optarg);
                if (qual_threshold < 0) { fprintf(stderr, "Quality threshold must be >= 0\n"); return EXIT_FAILURE; }
                break;
            case 'l':
                length_threshold = atoi(optarg);
                if (length_threshold < 0) { fprintf(stderr, "Length threshold must be >= 0\n"); return EXIT_FAILURE; }
                break;
            case 'x': no_fiveprime = 1; break;
            case 'n': trunc_n = 1; break;
            case 'g': gzip_output = 1; break;
            case 'z': quiet = 1; break;
            case 'd': debug = 1; break;
            case 'a': threads = atoi(optarg); threads_given = true; break;
            case 'b': batch_mib = atoi(optarg); break;
            case CHAR_MIN - 2: usage(EXIT_SUCCESS, NULL); break;
            case CHAR_MIN - 3: print_version_and_exit(); break;
            default: usage(EXIT_FAILURE, NULL); break;
        }
    }
    if (qualtype == -1) { usage(EXIT_FAILURE, "****Error: Quality type is required."); return EXIT_FAILURE; }
    if (!infn && !infnc) { usage(EXIT_FAILURE, "****Error: Must have either -f OR -c argument."); return EXIT_FAILURE; }
    return 0;
}

int Trim_Paired::trim_main() {
    ByteSource in0, in1;
    ByteSink o_main, o_mate2, o_single;
    ByteSink *outs[3] = {nullptr, nullptr, nullptr};
    int mode;
    const bool gz = gzip_output != 0;
    if (infnc) {   // interleaved input (reference init_streams, src/trim_paired.cpp:628-658)
        if (infn || infn2 || outfn || outfn2) { usage(EXIT_FAILURE, "****Error: Cannot have -f, -r, -o, or -p options with -c."); return EXIT_FAILURE; }
        if (outfnM && (outfnc || sfn)) { usage(EXIT_FAILURE, "****Error: Cannot have -m or -s options with -M."); return EXIT_FAILURE; }
        if (!outfnM && !outfnc) { usage(EXIT_FAILURE, "****Error: Must have -m or -M with -c."); return EXIT_FAILURE; }
        if (!in0.open(infnc)) { fprintf(stderr, "****Error: Could not open interleaved input file '%s'.\n\n", infnc); return EXIT_FAILURE; }
        const char *o = outfnM ? outfnM : outfnc;
        if (!o_main.open(o, gz)) { fprintf(stderr, "****Error: Could not open interleaved output file '%s'.\n\n", o); return EXIT_FAILURE; }
        outs[0] = &o_main;
        mode = outfnM ? SK_MODE_PE_INTER_M : SK_MODE_PE_INTER;
    } else {       // forward and reverse files (src/trim_paired.cpp:658-709)
        if (infn && (!infn2 || !outfn || !outfn2 || !sfn)) { usage(EXIT_FAILURE, "****Error: Using the -f option means you must have the -r, -o, -p, and -s options."); return EXIT_FAILURE; }
        if (infn && (infnc || outfnc || outfnM)) { usage(EXIT_FAILURE, "****Error: The -f option cannot be used in combination with -c, -m, or -M."); return EXIT_FAILURE; }
        if (!in0.open(infn)) { fprintf(stderr, "****Error: Could not open input file '%s'.\n\n", infn); return EXIT_FAILURE; }
        if (!in1.open(infn2)) { fprintf(stderr, "****Error: Could not open input file '%s'.\n\n", infn2); return EXIT_FAILURE; }
        if (!o_main.open(outfn, gz)) { fprintf(stderr, "****Error: Could not open output file '%s'.\n\n", outfn); return EXIT_FAILURE; }
        if (!o_mate2.open(outfn2, gz)) { fprintf(stderr, "****Error: Could not open output file '%s'.\n\n", outfn2); return EXIT_FAILURE; }
        outs[0] = &o_main;
        outs[1] = &o_mate2;
        mode = SK_MODE_PE_2FILE;
    }
    if (sfn) {
        if (!o_single.open(sfn, gz)) { fprintf(stderr, "****Error: Could not open single output file '%s'.\n\n", sfn); return EXIT_FAILURE; }
        outs[2] = &o_single;
    }
    Totals t;
    const int rc = run_device(mode, &in0, mode == SK_MODE_PE_2FILE ? &in1 : nullptr, outs, sfn != nullptr, t);
    const bool c0 = o_main.close(), c1 = o_mate2.close(), c2 = o_single.close();
    if (rc != EXIT_SUCCESS) return rc;
    if (!(c0 && c1 && c2)) { fprintf(stderr, "****Error: write failed\n\n"); return EXIT_FAILURE; }
    if (!quiet) {   // reference src/trim_paired.cpp:464-476, with the true record total
        const long long total = t.kept_p + t.kept_s1 + t.kept_s2 + t.discard_p + t.discard_s1 + t.discard_s2;
        if (infn && infn2) fprintf(stdout, "\nPE forward file: %s\nPE reverse file: %s\n", infn, infn2);
        if (infnc) fprintf(stdout, "\nPE interleaved file: %s\n", infnc);
        fprintf(stdout, "\nTotal input FastQ records: %lld (%lld pairs)\n", total, total / 2);
        fprintf(stdout, "\nFastQ paired records kept: %lld (%lld pairs)\n", t.kept_p, t.kept_p / 2);
        if (infnc) fprintf(stdout, "FastQ single records kept: %lld\n", t.kept_s1 + t.kept_s2);
        else fprintf(stdout, "FastQ single records kept: %lld (from PE1: %lld, from PE2: %lld)\n", t.kept_s1 + t.kept_s2, t.kept_s1, t.kept_s2);
        fprintf(stdout, "FastQ paired records discarded: %lld (%lld pairs)\n", t.discard_p, t.discard_p / 2);
        if (infnc) fprintf(stdout, "FastQ single records discarded: %lld\n\n", t.discard_s1 + t.discard_s2);
        else fprintf(stdout, "FastQ single records discarded: %lld (from PE1: %lld, from PE2: %lld)\n\n", t.discard_s1 + t.discard_s2, t.discard_s1, t.discard_s2);
    }
    if (debug)
        fprintf(stderr, "[sickle_b200] batches %lld (fused %lld), kernel %.3f ms | host: init %.3f s, read %.3f s, device wait %.3f s, "
                        "write wait %.3f s (writers busy %.3f s)\n",
                t.batches, t.fused_batches, t.kernel_ms, t.t_init, in0.read_seconds() + in1.read_seconds(), t.t_wait, t.t_write_wait,
                o_main.busy_seconds() + o_mate2.busy_seconds() + o_single.busy_seconds());
    return EXIT_SUCCESS;
}
