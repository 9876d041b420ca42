// io_tool.cpp -- copies a file through the CLI's I/O stages (ByteSource -> ByteSink), for the CPU
// tests of host/io.{h,cpp}:   io_tool <src> <dst> <read_chunk_bytes> <gzip_output 0|1>
// Prints "bgzf=<0|1> gzip=<0|1> bytes=<n>" on success; exit 1 on any read/write error.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "io.h"
#include "ref_batcher.h"
#include "unit_cutter.h"

// io_tool refbatch <src> <batch_len> <minlines> <cap>: prints the size of every reference batch
static int refbatch(int argc, char **argv) {
    if (argc != 6) { fprintf(stderr, "usage: io_tool refbatch <src> <batch_len> <minlines> <cap>\n"); return 2; }
    host::ByteSource in;
    if (!in.open(argv[2])) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    host::RefBatcher b(&in, atoll(argv[3]), atoi(argv[4]));
    std::vector<char> buf((size_t)strtoull(argv[5], nullptr, 10));
    const double t0 = host::now_s();
    unsigned long long total = 0;
    while (true) {
        const long long n = b.next(buf.data(), buf.size());
        if (n < 0) { printf("-1\n"); return 1; }
        if (n == 0) break;
        printf("%lld\n", n);
        total += (unsigned long long)n;
    }
    fprintf(stderr, "refbatch: %llu bytes in %.3f s\n", total, host::now_s() - t0);
    return 0;
}

// io_tool units <cap> <lines_per_unit> <src0> [<src1>]: the whole-unit batches the multi-GPU driver
// would dispatch (host/unit_cutter.h), one line "<bytes0> <bytes1> <units>" per batch.  With two sources
// (paired files) a batch holds the same number of records of each, lines_per_unit applies to src0
// (src1 is always 4).  Exit 3: a full buffer without a complete unit.
static int units(int argc, char **argv) {
    if (argc != 5 && argc != 6) { fprintf(stderr, "usage: io_tool units <cap> <lines_per_unit> <src0> [<src1>]\n"); return 2; }
    const unsigned long long cap = strtoull(argv[2], nullptr, 10);
    const bool two = argc == 6;
    host::ByteSource in0, in1;
    if (!in0.open(argv[4]) || (two && !in1.open(argv[5]))) { fprintf(stderr, "cannot open input\n"); return 1; }
    host::UnitStream s0(&in0, atoi(argv[3])), s1(two ? &in1 : &in0, 4);
    // two buffers per input: the carried bytes stay in the previous buffer until the next fill()
    std::vector<char> buf[2][2];
    for (auto &a : buf) for (auto &b : a) b.resize((size_t)cap);
    for (int k = 0;; k ^= 1) {
        if (!s0.fill(buf[0][k].data(), cap) || (two && !s1.fill(buf[1][k].data(), cap))) { fprintf(stderr, "read failed\n"); return 1; }
        const unsigned long long u = two ? std::min(s0.units(), s1.units()) : s0.units();
        if (u == 0) {
            if ((s0.units() == 0 && s0.full() && !s0.eof()) || (two && s1.units() == 0 && s1.full() && !s1.eof())) return 3;
            break;
        }
        const unsigned long long n0 = s0.cut(u), n1 = two ? s1.cut(u) : 0;
        printf("%llu %llu %llu\n", n0, n1, u);
    }
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 2 && std::string(argv[1]) == "refbatch") return refbatch(argc, argv);
    if (argc >= 2 && std::string(argv[1]) == "units") return units(argc, argv);
    if (argc != 5) { fprintf(stderr, "usage: io_tool <src> <dst> <chunk> <gzip 0|1>\n"); return 2; }
    const unsigned long long chunk = strtoull(argv[3], nullptr, 10);
    host::ByteSource in;
    if (!in.open(argv[1])) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    host::ByteSink out;
    if (!out.open(argv[2], atoi(argv[4]) != 0)) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    // two buffers in flight, like the slots of the trimmer
    std::vector<char> buf[2] = {std::vector<char>(chunk), std::vector<char>(chunk)};
    unsigned long long ticket[2] = {0, 0}, total = 0;
    for (int k = 0;; k ^= 1) {
        if (ticket[k] && !out.wait(ticket[k])) { fprintf(stderr, "write failed\n"); return 1; }
        const long long r = in.read(buf[k].data(), chunk);
        if (r < 0) { fprintf(stderr, "read failed\n"); return 1; }
        if (r == 0) break;
        ticket[k] = out.write_async(buf[k].data(), (unsigned long long)r);
        total += (unsigned long long)r;
        if ((unsigned long long)r < chunk) break;
    }
    if (!out.close()) { fprintf(stderr, "write failed\n"); return 1; }
    printf("bgzf=%d gzip=%d bytes=%llu\n", in.bgzf() ? 1 : 0, in.gzip() ? 1 : 0, total);
    return 0;
}
