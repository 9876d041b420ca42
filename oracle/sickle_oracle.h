/*
 * sickle_oracle.h -- CPU restatement of Parallel Sickle's per-read trimming hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, link, load or execute it, and only as the checker.  The product
 * (sickle_b200/, include/) never includes this header and has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against (a) md5 fixtures produced by the unmodified reference binary
 * (oracle/_ref/sickle, built from /root/reference/src by oracle/Makefile) on the
 * reference's bundled test/ FASTQ files, and (b) reference outputs on synthetic
 * inputs committed under tests/golden/ with the generating script.
 * Exception: PE -M ("N-record" mode) has no reference implementation in this fork
 *  -- that mode is "parity unpinned" and follows README.md:116-120.
 *
 * Each function cites the reference file:line it follows (paths under
 * /root/reference/src/).
 */
#ifndef SICKLE_ORACLE_H
#define SICKLE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* quality_type, sickle.h:61-66 (PHRED=0 unused by the CLI) */
enum { SO_SANGER = 1, SO_SOLEXA = 2, SO_ILLUMINA = 3 };

/* error kinds; 1..5 = FQEntry::validate (FQEntry.cpp:53-97), 6 = get_quality_num (trim.cpp:129-137) */
enum {
    SO_OK = 0,
    SO_ERR_ID_SHORT = 1,      /* FQEntry.cpp:55  name.length() <= 1        */
    SO_ERR_ID_CHAR = 2,       /* FQEntry.cpp:66  name[0] != '@'            */
    SO_ERR_SEQ_EMPTY = 3,     /* FQEntry.cpp:76                            */
    SO_ERR_QUAL_EMPTY = 4,    /* FQEntry.cpp:82                            */
    SO_ERR_LEN_MISMATCH = 5,  /* FQEntry.cpp:88                            */
    SO_ERR_QUAL_RANGE = 6,    /* trim.cpp:129                              */
    SO_ERR_BATCH_MISMATCH = 7 /* trim_paired.cpp:335-338 (pe -f/-r only)   */
};

enum { SO_MODE_SE = 0, SO_MODE_PE_2FILE = 1, SO_MODE_PE_INTER = 2, SO_MODE_PE_INTER_M = 3 };

typedef struct {
    int qualtype;      /* SO_SANGER / SO_SOLEXA / SO_ILLUMINA                */
    int qual_threshold;   /* -q, default 20 (trim_single.cpp:70)             */
    int length_threshold; /* -l, default 20 (trim_single.cpp:69)             */
    int no_fiveprime;  /* -x                                                 */
    int trunc_n;       /* -n                                                 */
} so_params;

typedef struct {
    int five;          /* cutsites.five_prime_cut  (sickle.h:93-96)          */
    int three;         /* cutsites.three_prime_cut; < 0 => discard           */
} so_cut;

typedef struct {
    int kind;          /* SO_OK or SO_ERR_*                                  */
    int64_t record;    /* 0-based record number within its input file        */
    int file;          /* 0 = first input, 1 = second input (pe -r)          */
    int position;      /* 0-based byte position in the quality string        */
    int byte;          /* offending quality byte as (signed char) value      */
} so_error;

typedef struct {
    int64_t kept, discard;                         /* SE  (trim_single.cpp:391,397) */
    int64_t kept_p, discard_p;                     /* PE  (trim_paired.cpp:551,566) */
    int64_t kept_s1, kept_s2, discard_s1, discard_s2;
    int64_t records_in[2];                         /* complete records consumed per input */
    int64_t n_batches;
} so_counters;

/* One read: trim.cpp:3-116.  seq/qual need not be NUL-terminated.  Returns SO_OK or
 * SO_ERR_QUAL_RANGE (err->position / err->byte filled).  `visited` (optional) receives
 * the number of leading quality bytes the scalar loop range-checked. */
int so_sliding_window(const char *seq, size_t seq_len, const char *qual, size_t qual_len,
                      const so_params *p, so_cut *cut, so_error *err, int *visited);

/* Reference batch geometry: trim_single.cpp:194-211 / trim_paired.cpp:246-263. */
int64_t so_recommended_batch_len(int64_t file_size, int64_t b_mib, int paired);

/*
 * Whole-file drivers.  `threads` = the reference's -a N (output order policy,
 * trim_single.cpp:263,273-274 / trim_paired.cpp:349,388,403); `batch_len` = the value the
 * reference would compute (so_recommended_batch_len) -- only matters when threads > 1.
 * Outputs are written into caller buffers of capacity >= input size (+ n2 for out[1] etc.);
 * out_len[] receives the byte counts.
 *   SE:           out[0]
 *   PE_2FILE:     out[0] = -o, out[1] = -p, out[2] = -s
 *   PE_INTER:     out[0] = -m, out[2] = -s (pass has_singles = 0 to drop singles, trim_paired.cpp:601)
 *   PE_INTER_M:   out[0] = -M
 * Returns SO_OK or the first error (err filled); outputs then hold what the reference would
 * have emitted for the batches completed before the error.
 */
int so_run(int mode, const so_params *p, int threads, int64_t batch_len, int has_singles,
           const char *in1, size_t n1, const char *in2, size_t n2,
           char *out[3], size_t out_cap[3], size_t out_len[3],
           so_counters *ctr, so_error *err);

#ifdef __cplusplus
}
#endif
#endif
