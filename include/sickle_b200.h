/*
 * sickle_b200.h -- C ABI of the B200-native trimming hot path of Parallel Sickle.
 *
 * The reference (pentalpha/sickle) has no plugin / FFI interface: its only seam is the C++ class
 * Abstract_Trimmer (src/trim.h:8-38) whose trim_main() loops
 *     GZReader::get_batch_buffering_lines()  (src/GZReader.cpp:29-41,59-132)
 *  -> FQEntry(...) + validate()              (src/FQEntry.cpp:8-18,53-97)
 *  -> sliding_window() per read              (src/trim.cpp:3-116, get_quality_num :118-140)
 *  -> output_single() / output_paired()      (src/trim_single.cpp:374-428, src/trim_paired.cpp:506-624)
 * This header is the boundary a maintainer would put under that loop: the host keeps the CLI,
 * file / zlib reading and file writing; everything from "bytes of one batch in a pinned buffer"
 * to "bytes of the trimmed output streams + counters in pinned buffers" happens behind these
 * entry points, on one B200 per context.  INTEGRATION.md shows the reference-side call sites.
 *
 * Plain C types only.  No CPU fallback: every entry point fails (negative return, text in
 * sk_last_error()) when no CUDA device is usable.
 *
 * Threading: one context per GPU; calls on one context must come from one host thread at a
 * time.  Different contexts (different GPUs) may be driven from different threads.
 */
#ifndef SICKLE_B200_H
#define SICKLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_ABI_VERSION 1

/* quality_type of the reference, src/sickle.h:61-66 (PHRED = 0 is not reachable from its CLI) */
enum { SK_QUAL_SANGER = 1, SK_QUAL_SOLEXA = 2, SK_QUAL_ILLUMINA = 3 };

/* which reference driver the batch belongs to */
enum {
    SK_MODE_SE = 0,         /* sickle se                       (src/trim_single.cpp)              */
    SK_MODE_PE_2FILE = 1,   /* sickle pe -f/-r -o/-p/-s        (src/trim_paired.cpp, two readers)  */
    SK_MODE_PE_INTER = 2,   /* sickle pe -c -m [-s]            (interleaved reader)                */
    SK_MODE_PE_INTER_M = 3  /* sickle pe -c -M  (README.md:116-120; absent from the fork's code)   */
};

/* output stream indices in sk_result.out[] */
enum {
    SK_OUT_MAIN = 0,   /* se: -o      pe 2-file: -o (mate 1)      interleaved: -m / -M */
    SK_OUT_MATE2 = 1,  /*             pe 2-file: -p (mate 2)                             */
    SK_OUT_SINGLES = 2 /*             pe: -s                                             */
};

/* data errors the reference reports with exit(1); 1..5 = FQEntry::validate (src/FQEntry.cpp:55-94),
 * 6 = get_quality_num (src/trim.cpp:129-137) */
enum {
    SK_DATA_OK = 0,
    SK_DATA_ID_SHORT = 1,
    SK_DATA_ID_CHAR = 2,
    SK_DATA_SEQ_EMPTY = 3,
    SK_DATA_QUAL_EMPTY = 4,
    SK_DATA_LEN_MISMATCH = 5,
    SK_DATA_QUAL_RANGE = 6
};

/* return codes */
enum {
    SK_OK = 0,
    SK_E_ARG = -1,       /* bad argument / state                                          */
    SK_E_CUDA = -2,      /* CUDA runtime failure (fatal for the context)                  */
    SK_E_NOMEM = -3,
    SK_E_CAPACITY = -4   /* batch has more lines than the slot's line index can hold      */
};

/* Mirrors the option members of Abstract_Trimmer (src/trim.h:18-27). */
typedef struct sk_params {
    int32_t qualtype;          /* -t : SK_QUAL_*                                            */
    int32_t qual_threshold;    /* -q : default 20 (src/trim_single.cpp:70)                  */
    int32_t length_threshold;  /* -l : default 20 (src/trim_single.cpp:69)                  */
    int32_t no_fiveprime;      /* -x                                                        */
    int32_t trunc_n;           /* -n                                                        */
    int32_t mode;              /* SK_MODE_*                                                 */
    int32_t emulate_threads;   /* reference -a N output order inside a batch; <= 1: input order
                                  (src/trim_single.cpp:263,273-274; src/trim_paired.cpp:349,388,403) */
    int32_t has_singles;       /* pe: a -s file was given (src/trim_paired.cpp:601,609)     */
} sk_params;

typedef struct sk_error_info {
    int32_t kind;       /* SK_DATA_*                                                      */
    int32_t file;       /* 0 = first input buffer, 1 = second                             */
    int64_t record;     /* 0-based record number inside that input buffer of this batch  */
    int32_t position;   /* SK_DATA_QUAL_RANGE: 0-based index into the quality string      */
    int32_t byte;       /* SK_DATA_QUAL_RANGE: the quality byte as a signed char value    */
    /* byte ranges of the offending record's four lines, as offsets into the input buffer `file`
       (so the host can print the reference's messages without re-parsing): */
    uint64_t line_off[4];
    uint64_t line_len[4];
} sk_error_info;

typedef struct sk_result {
    const char *out[3];      /* pinned host memory, valid until the slot is submitted again   */
    uint64_t out_bytes[3];
    uint64_t consumed[2];    /* bytes of each input, counted from `start`, that formed complete
                                records (pairs); the caller carries the rest into the next batch
                                (the reference's last_remainder, src/GZReader.cpp:104-129)     */
    uint64_t records[2];     /* complete records consumed per input                           */
    int64_t kept, discard;                          /* se (src/trim_single.cpp:391,397)       */
    int64_t kept_p, discard_p;                      /* pe (src/trim_paired.cpp:551,566)       */
    int64_t kept_s1, kept_s2, discard_s1, discard_s2;
    sk_error_info error;     /* error.kind != 0: first data error of the batch; outputs invalid */
    float kernel_ms;         /* device time of the batch's kernels (CUDA events)              */
    float stage_ms[4];       /* of which: K1 line index, K2 trim+route, K3 emit, summary      */
    uint32_t kernel_launches;/* kernels launched for the batch                                */
    uint32_t fused;          /* 1: the single-pass fused kernel produced the batch; 0: K1/K2/K3 */
} sk_result;

typedef struct sk_ctx sk_ctx;

/* Library / device probes.  sk_device_count() < 0 means CUDA is unusable. */
int sk_abi_version(void);
int sk_device_count(void);
const char *sk_last_error(void);

/*
 * Create a context on `device` with `n_slots` pipeline slots.  Each slot owns pinned host input
 * buffer(s) of `slot_bytes` (two when mode is SK_MODE_PE_2FILE), pinned output buffers, and the
 * matching device buffers + line index.  slot_bytes <= 2^31 - 4096.  n_slots == 0 creates a
 * context with device scratch only (for sk_trim_device).
 */
sk_ctx *sk_create(int device, uint64_t slot_bytes, int n_slots, const sk_params *params);
void sk_destroy(sk_ctx *ctx);

/* Pinned host input buffer of a slot (which = 0 or 1), capacity sk_slot_bytes().  The caller puts
 * FASTQ bytes in it.  Replaces the per-line `new char[]` + strncpy of GZReader::read_lines
 * (src/GZReader.cpp:76-92). */
char *sk_in_buffer(sk_ctx *ctx, int slot, int which);
uint64_t sk_slot_bytes(const sk_ctx *ctx);

/*
 * Optional early upload: start the asynchronous H2D copy of buffer bytes [offset, offset+nbytes)
 * before the batch's first byte is known.  A pipelined reader fills the bulk of the next slot at
 * some headroom offset and uploads it while the previous batch is still running; when that batch
 * reports `consumed`, the reader copies the carried-over tail (the reference's last_remainder,
 * src/GZReader.cpp:104-129) in front of the bulk and calls sk_submit.  At most one sk_upload range
 * per input and batch.
 */
int sk_upload(sk_ctx *ctx, int slot, int which, uint64_t offset, uint64_t nbytes);

/*
 * Asynchronously run one batch made of buffer bytes [start0, end0) of input 0 (and [start1, end1)
 * of input 1 for SK_MODE_PE_2FILE; pass 0, 0 otherwise).  data[start] must be the first byte of a
 * record.  Bytes of the range not covered by an earlier sk_upload are uploaded now.  Then:
 * K1 line index (Batch/FQEntry split), K2 sliding_window + routing scan, K3 formatting, and the
 * device-side summary.  Returns immediately.  Only whole lines are looked at: bytes after the last
 * '\n', and trailing lines that do not complete a record (pair), are left unconsumed.
 */
int sk_submit(sk_ctx *ctx, int slot, uint64_t start0, uint64_t end0, uint64_t start1, uint64_t end1);

/* Block until the slot's batch is done, copy the output streams to the slot's pinned output
 * buffers (exact sizes) and fill `res`. */
int sk_wait(sk_ctx *ctx, int slot, sk_result *res);

/*
 * Device-resident variant (kernel-only path; also what an embedding GPU pipeline would call):
 * inputs and outputs are device pointers owned by the caller, `stream` is a cudaStream_t (NULL =
 * the context's stream of that slot).  in0/in1 need 16-byte alignment and 16 readable padding
 * bytes after n0/n1.  Outputs stay on the device; sk_result.out[] is NULL.  Uses the scratch of
 * `slot` (0 when the context was created with n_slots == 0).
 */
int sk_trim_device(sk_ctx *ctx, int slot, const void *in0, uint64_t n0, const void *in1, uint64_t n1,
                   void *const out[3], const uint64_t out_cap[3], void *stream);
/* Synchronise `stream` and fetch the summary of the last sk_trim_device on `slot`. */
int sk_result_device(sk_ctx *ctx, int slot, void *stream, sk_result *res);

#ifdef __cplusplus
}
#endif
#endif /* SICKLE_B200_H */
