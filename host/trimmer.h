// trimmer.h -- host side of the drop-in `sickle se` / `sickle pe` command line.
//
// Mirrors the reference's operator interface for this path: Abstract_Trimmer with parse_args(),
// trim_main() and usage() (reference src/trim.h:8-38), implemented by Trim_Single
// (src/trim_single.{h,cpp}) and Trim_Paired (src/trim_paired.{h,cpp}).  Same option letters, same
// defaults, same messages and exit codes; what differs is trim_main(): instead of splitting lines on
// the host and running sliding_window() in std::threads, it streams whole byte batches through the
// C ABI of include/sickle_b200.h (one B200 per context, pinned slots, H2D / kernels / D2H overlapped).
#ifndef SICKLE_B200_HOST_TRIMMER_H
#define SICKLE_B200_HOST_TRIMMER_H

#include <cstdint>
#include <string>
#include <vector>

#include "io.h"
#include "sickle_b200.h"

#ifndef PROGRAM_NAME
#define PROGRAM_NAME "sickle"
#endif
#ifndef SICKLE_VERSION
#define SICKLE_VERSION 1.33
#endif

namespace host {

// Set by `sickle batch`: several commands run in this process, contexts are kept between them.
extern bool batch_mode;

struct Totals {
    long long kept = 0, discard = 0;
    long long kept_p = 0, discard_p = 0, kept_s1 = 0, kept_s2 = 0, discard_s1 = 0, discard_s2 = 0;
    long long records[2] = {0, 0};
    double kernel_ms = 0;
    double t_init = 0, t_wait = 0, t_write_wait = 0, t_teardown = 0;   // host stage times (-d)
    long long batches = 0, fused_batches = 0;
};

}  // namespace host

class Abstract_Trimmer {
public:
    virtual ~Abstract_Trimmer() {}
    virtual int parse_args(int argc, char *argv[]) = 0;
    virtual int trim_main() = 0;
    virtual void usage(int status, char const *msg) = 0;

protected:
    // option state, same meaning as the reference members (src/trim.h:18-27)
    int qualtype = -1;
    int length_threshold = 20;
    int qual_threshold = 20;
    int no_fiveprime = 0;
    int trunc_n = 0;
    int debug = 0;
    int threads = 1;          // -a: reference output order to reproduce (1 = input order)
    bool threads_given = false;
    long long batch_mib = 512;   // -b
    int quiet = 0;
    int gzip_output = 0;
    char *infn = nullptr;
    char *outfn = nullptr;

    // Run one input stream (se, or interleaved pe) / two input streams through the device.
    // outs[k] may be null (stream not written).  Returns the process exit code.
    int run_device(int mode, host::ByteSource *in0, host::ByteSource *in1, host::ByteSink *outs[3],
                   bool has_singles, host::Totals &tot);
    // The same over several GPUs (independent whole-record batches dealt to one context per device).
    int run_devices(const std::vector<int> &devices, const sk_params &p, host::ByteSource *in0, host::ByteSource *in1,
                    host::ByteSink *outs[3], host::Totals &tot);
    int report_data_error(const sk_result &r, const char *buf0, const char *buf1);
};

class Trim_Single : public Abstract_Trimmer {
public:
    int parse_args(int argc, char *argv[]) override;
    int trim_main() override;
    void usage(int status, char const *msg) override;
};

class Trim_Paired : public Abstract_Trimmer {
public:
    int parse_args(int argc, char *argv[]) override;
    int trim_main() override;
    void usage(int status, char const *msg) override;

private:
    char *infn2 = nullptr, *infnc = nullptr;
    char *outfn2 = nullptr, *outfnc = nullptr, *sfn = nullptr, *outfnM = nullptr;
};

#endif
