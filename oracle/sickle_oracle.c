/*
 * sickle_oracle.c -- plain-C restatement of the reference's hot path (see sickle_oracle.h).
 * TEST INFRASTRUCTURE ONLY: never linked into, loaded by, or called from the product path.
 *
 * It deliberately keeps the reference's *scalar* shape (one window slid base by base, with
 * a double-precision average) so that it is an independent check on the integer,
 * data-parallel formulation the CUDA kernels use.
 */
#include "sickle_oracle.h"

#include <stdlib.h>
#include <string.h>

/* quality_constants, sickle.h:85-91: {offset, min, max} */
static const int QC[4][3] = {
    {0, 4, 60}, {33, 33, 126}, {64, 58, 112}, {64, 64, 110}};

/* get_quality_num, trim.cpp:118-140.  qualchar is a (signed) char in the reference; an
 * out-of-range value makes the reference print six lines to stderr and exit(1). */
static int quality_num(const char *qual, int pos, const so_params *p, so_error *err, int *bad) {
    int v = (int)(signed char)qual[pos];
    if (v < QC[p->qualtype][1] || v > QC[p->qualtype][2]) {
        if (!*bad && err) {
            err->kind = SO_ERR_QUAL_RANGE;
            err->position = pos;
            err->byte = v;
        }
        *bad = 1;
    }
    return v - QC[p->qualtype][0];
}

/* sliding_window, trim.cpp:3-116 */
int so_sliding_window(const char *seq, size_t seq_len, const char *qual, size_t qual_len,
                      const so_params *p, so_cut *cut, so_error *err, int *visited) {
    int window_size = (int)(0.1 * (double)seq_len);           /* trim.cpp:8  */
    int i, j;
    int window_start = 0, window_total = 0;
    int three_prime_cut = (int)seq_len;                        /* trim.cpp:13 */
    int five_prime_cut = 0, found_five_prime = 0;
    int bad = 0;
    int nvisited = 0;
    double window_avg;

    if (visited) *visited = 0;
    if (seq_len < (size_t)p->length_threshold) {               /* trim.cpp:21-26 */
        cut->five = -1;
        cut->three = -1;
        return SO_OK;
    }
    if (window_size == 0) window_size = (int)seq_len;          /* trim.cpp:30 */

    for (i = 0; i < window_size; i++) {                        /* trim.cpp:31-33 */
        window_total += quality_num(qual, i, p, err, &bad);
        if (bad) return SO_ERR_QUAL_RANGE;
        nvisited = i + 1;
    }
    for (i = 0; (size_t)i <= qual_len - (size_t)window_size; i++) { /* trim.cpp:34 */
        window_avg = (double)window_total / (double)window_size;    /* trim.cpp:36 */

        if (p->no_fiveprime == 0 && found_five_prime == 0 &&
            window_avg >= (double)p->qual_threshold) {              /* trim.cpp:42 */
            for (j = window_start; j < window_start + window_size; j++) { /* :46-51 */
                if (quality_num(qual, j, p, err, &bad) >= p->qual_threshold) {
                    five_prime_cut = j;
                    break;
                }
            }
            found_five_prime = 1;
        }
        if ((window_avg < (double)p->qual_threshold ||
             (size_t)(window_start + window_size) > qual_len) &&
            (found_five_prime == 1 || p->no_fiveprime)) {           /* trim.cpp:61-62 */
            for (j = window_start; j < window_start + window_size; j++) { /* :65-70 */
                if (quality_num(qual, j, p, err, &bad) < p->qual_threshold) {
                    three_prime_cut = j;
                    break;
                }
            }
            break;
        }
        window_total -= quality_num(qual, window_start, p, err, &bad); /* trim.cpp:76 */
        if ((size_t)(window_start + window_size) < qual_len) {      /* trim.cpp:77-79 */
            window_total += quality_num(qual, window_start + window_size, p, err, &bad);
            if (bad) return SO_ERR_QUAL_RANGE;
            nvisited = window_start + window_size + 1;
        }
        window_start++;
    }
    if (visited) *visited = nvisited;

    /* -n, trim.cpp:86-98: lowercase 'n' wins; an uppercase-only 'N' leaves npos = string::npos,
     * so (int)(npos - 1) == -2 (reproduced on purpose). */
    if (p->trunc_n) {
        const char *pn = seq_len ? (const char *)memchr(seq, 'n', seq_len) : NULL;
        const char *pN = seq_len ? (const char *)memchr(seq, 'N', seq_len) : NULL;
        if (pn) three_prime_cut = (int)(pn - seq) - 1;
        else if (pN) three_prime_cut = -2;
    }
    if ((found_five_prime == 0 && !p->no_fiveprime) ||
        (three_prime_cut - five_prime_cut < p->length_threshold)) {  /* trim.cpp:103 */
        three_prime_cut = -1;
        five_prime_cut = -1;
    }
    cut->five = five_prime_cut;
    cut->three = three_prime_cut;
    return SO_OK;
}

/* trim_single.cpp:194-211, trim_paired.cpp:246-263 */
int64_t so_recommended_batch_len(int64_t file_size, int64_t b_mib, int paired) {
    int64_t max = (int64_t)(uint32_t)(int32_t)(b_mib * 1024 * 1024);
    int64_t rec = file_size / 8;
    if (paired) max /= 2;
    if (rec < 20) return 20;
    if (rec > max) return max;
    return rec;
}

/* ---- line reader: GZReader::read_lines, GZReader.cpp:59-132 ------------------------- */
typedef struct { const char *p; size_t len; } line_t;
typedef struct { line_t *v; size_t n, cap; } linevec;

static void lv_push(linevec *lv, line_t l) {
    if (lv->n == lv->cap) {
        lv->cap = lv->cap ? lv->cap * 2 : 1024;
        lv->v = (line_t *)realloc(lv->v, lv->cap * sizeof(line_t));
    }
    lv->v[lv->n++] = l;
}

typedef struct {
    const char *buf; size_t n, pos;
    int eof, minlines;
    int64_t batch_len;
    line_t carry[8]; int ncarry;
} reader_t;

/* Fills `out` with the next batch's lines (multiple of minlines); returns their count. */
static size_t read_batch(reader_t *r, linevec *out) {
    int64_t remaining = r->batch_len;
    out->n = 0;
    if (r->eof) return 0;                                   /* GZReader.cpp:31 */
    for (int i = 0; i < r->ncarry; i++) {                   /* GZReader.cpp:68-75 */
        remaining -= (int64_t)r->carry[i].len;
        lv_push(out, r->carry[i]);
    }
    r->ncarry = 0;
    do {                                                    /* GZReader.cpp:76-92 */
        if (r->pos >= r->n) { r->eof = 1; break; }          /* gzgets -> NULL */
        const char *s = r->buf + r->pos;
        const char *nl = (const char *)memchr(s, '\n', r->n - r->pos);
        size_t got = nl ? (size_t)(nl - s) + 1 : r->n - r->pos; /* chars gzgets returned */
        line_t l;
        l.p = s;
        l.len = got - 1;      /* content = all but the last char ('\n', or a real char at an
                                 unterminated EOF line: GZReader.cpp:81-88) */
        remaining -= (int64_t)l.len;
        r->pos += got;
        lv_push(out, l);
    } while (remaining > 0);
    size_t extra = out->n % (size_t)r->minlines;            /* GZReader.cpp:104-129 */
    for (size_t i = 0; i < extra; i++) r->carry[i] = out->v[out->n - extra + i];
    r->ncarry = (int)extra;
    out->n -= extra;
    return out->n;
}

/* FQEntry::validate, FQEntry.cpp:53-97 (line 3 is not inspected) */
static int validate(const line_t *l) {
    if (l[0].len <= 1) return SO_ERR_ID_SHORT;
    if (l[0].p[0] != '@') return SO_ERR_ID_CHAR;
    if (l[1].len < 1) return SO_ERR_SEQ_EMPTY;
    if (l[3].len < 1) return SO_ERR_QUAL_EMPTY;
    if (l[3].len != l[1].len) return SO_ERR_LEN_MISMATCH;
    return SO_OK;
}

typedef struct { char *p; size_t len, cap; int overflow; } outbuf;

static void ob_put(outbuf *o, const char *s, size_t n) {
    if (o->len + n > o->cap) { o->overflow = 1; return; }
    memcpy(o->p + o->len, s, n);
    o->len += n;
}
static void ob_nl(outbuf *o) { ob_put(o, "\n", 1); }

/* trim_single.cpp:393-396, trim_paired.cpp:506-513: line 3 echoed verbatim */
static void emit_record(outbuf *o, const line_t *l, so_cut c) {
    size_t n = (size_t)(c.three - c.five);
    ob_put(o, l[0].p, l[0].len); ob_nl(o);
    ob_put(o, l[1].p + c.five, n); ob_nl(o);
    ob_put(o, l[2].p, l[2].len); ob_nl(o);
    ob_put(o, l[3].p + c.five, n); ob_nl(o);
}

/* -M "N record" (README.md:116-120, sickle.xml:204-206; unpinned):
 * name verbatim, sequence "N", line 3 verbatim, quality = the type's Q_MIN char. */
static void emit_n_record(outbuf *o, const line_t *l, int qualtype) {
    char q = (char)QC[qualtype][1];
    ob_put(o, l[0].p, l[0].len); ob_nl(o);
    ob_put(o, "N", 1); ob_nl(o);
    ob_put(o, l[2].p, l[2].len); ob_nl(o);
    ob_put(o, &q, 1); ob_nl(o);
}

int so_run(int mode, const so_params *p, int threads, int64_t batch_len, int has_singles,
           const char *in1, size_t n1, const char *in2, size_t n2,
           char *out[3], size_t out_cap[3], size_t out_len[3],
           so_counters *ctr, so_error *err) {
    reader_t r1, r2;
    linevec b1 = {0, 0, 0}, b2 = {0, 0, 0};
    outbuf ob[3];
    so_cut *cuts1 = NULL, *cuts2 = NULL;
    size_t cuts_cap = 0;
    int rc = SO_OK;
    const int paired = (mode != SO_MODE_SE);
    const int inter = (mode == SO_MODE_PE_INTER || mode == SO_MODE_PE_INTER_M);
    so_counters c;
    so_error e;

    memset(&c, 0, sizeof c);
    memset(&e, 0, sizeof e);
    if (threads < 1) threads = 1;
    for (int s = 0; s < 3; s++) {
        ob[s].p = out ? out[s] : NULL;
        ob[s].cap = (out && out[s]) ? out_cap[s] : 0;
        ob[s].len = 0;
        ob[s].overflow = 0;
    }
    memset(&r1, 0, sizeof r1);
    memset(&r2, 0, sizeof r2);
    r1.buf = in1; r1.n = n1; r1.batch_len = batch_len;
    r1.minlines = inter ? 8 : 4;                             /* GZReader.cpp:7-11 */
    r2.buf = in2; r2.n = n2; r2.batch_len = batch_len; r2.minlines = 4;

    for (;;) {
        size_t nl1 = read_batch(&r1, &b1);                   /* trim_single.cpp:245 */
        if (nl1 == 0) break;
        if (mode == SO_MODE_PE_2FILE) {                      /* trim_paired.cpp:326-338 */
            size_t nl2 = read_batch(&r2, &b2);
            if (nl2 == 0) break;
            if (nl2 != nl1) { rc = SO_ERR_BATCH_MISMATCH; e.kind = rc; break; }
        }
        c.n_batches++;

        /* units of this batch: SE record, or PE pair */
        size_t nunits = paired ? (inter ? nl1 / 8 : nl1 / 4) : nl1 / 4;
        if (nunits > cuts_cap) {
            cuts_cap = nunits;
            cuts1 = (so_cut *)realloc(cuts1, cuts_cap * sizeof(so_cut));
            cuts2 = (so_cut *)realloc(cuts2, cuts_cap * sizeof(so_cut));
        }
#define REC1(k) (inter ? &b1.v[8 * (k)] : &b1.v[4 * (k)])
#define REC2(k) (inter ? &b1.v[8 * (k) + 4] : &b2.v[4 * (k)])

        /* dealing loop: every record of the batch is validated before any is trimmed
         * (trim_single.cpp:265-298, trim_paired.cpp:350-404) */
        size_t ndealt = 0;
        int64_t chars_read = 0;
        for (size_t k = 0; k < nunits; k++) {
            if (paired && chars_read > batch_len) break;     /* trim_paired.cpp:352-358 */
            int v = validate(REC1(k));
            if (v) {
                rc = v; e.kind = v; e.file = 0;
                e.record = c.records_in[0] + (int64_t)(inter ? 2 * k : k);
                break;
            }
            if (paired) {
                v = validate(REC2(k));
                if (v) {
                    rc = v; e.kind = v; e.file = inter ? 0 : 1;
                    e.record = inter ? c.records_in[0] + (int64_t)(2 * k + 1)
                                     : c.records_in[1] + (int64_t)k;
                    break;
                }
            }
            chars_read += (int64_t)REC1(k)[1].len;
            ndealt++;
        }
        if (rc) break;

        /* processing_thread: trim_single.cpp:357-372, trim_paired.cpp:483-504 */
        for (size_t k = 0; k < ndealt && !rc; k++) {
            const line_t *a = REC1(k);
            rc = so_sliding_window(a[1].p, a[1].len, a[3].p, a[3].len, p, &cuts1[k], &e, NULL);
            if (rc) {
                e.file = 0;
                e.record = c.records_in[0] + (int64_t)(inter ? 2 * k : k);
                break;
            }
            if (paired) {
                const line_t *b = REC2(k);
                rc = so_sliding_window(b[1].p, b[1].len, b[3].p, b[3].len, p, &cuts2[k], &e, NULL);
                if (rc) {
                    e.file = inter ? 0 : 1;
                    e.record = inter ? c.records_in[0] + (int64_t)(2 * k + 1)
                                     : c.records_in[1] + (int64_t)k;
                }
            }
        }
        if (rc) break;

        /* output_single / output_paired: queue 0..N-1, ascending inside a queue */
        for (int q = 0; q < threads; q++) {
            /* SE: record k -> queue (k+1)%N (trim_single.cpp:263,273-274);
             * PE: pair   k -> queue  k   %N (trim_paired.cpp:349,388,403) */
            size_t first = paired ? (size_t)q : (size_t)((q + threads - 1) % threads);
            for (size_t k = first; k < ndealt; k += (size_t)threads) {
                const line_t *a = REC1(k);
                int k1 = cuts1[k].three >= 0;                 /* trim_single.cpp:368 */
                if (!paired) {
                    if (k1) { emit_record(&ob[0], a, cuts1[k]); c.kept++; }
                    else c.discard++;
                    continue;
                }
                const line_t *b = REC2(k);
                int k2 = cuts2[k].three >= 0;
                if (k1 && k2) {                               /* trim_paired.cpp:543-551 */
                    emit_record(&ob[0], a, cuts1[k]);
                    emit_record(inter ? &ob[0] : &ob[1], b, cuts2[k]);
                    c.kept_p += 2;
                } else if (k1 || k2) {                        /* trim_paired.cpp:552-563 */
                    if (mode == SO_MODE_PE_INTER_M) {
                        if (k1) { emit_record(&ob[0], a, cuts1[k]); emit_n_record(&ob[0], b, p->qualtype); }
                        else    { emit_n_record(&ob[0], a, p->qualtype); emit_record(&ob[0], b, cuts2[k]); }
                    } else if (has_singles) {                 /* trim_paired.cpp:601,609 */
                        if (k1) emit_record(&ob[2], a, cuts1[k]);
                        else    emit_record(&ob[2], b, cuts2[k]);
                    }
                    if (k1) { c.kept_s1++; c.discard_s2++; }
                    else    { c.kept_s2++; c.discard_s1++; }
                } else {                                      /* trim_paired.cpp:564-567 */
                    if (mode == SO_MODE_PE_INTER_M) {
                        emit_n_record(&ob[0], a, p->qualtype);
                        emit_n_record(&ob[0], b, p->qualtype);
                    }
                    c.discard_p += 2;
                }
            }
        }
        c.records_in[0] += (int64_t)(inter ? 2 * ndealt : ndealt);
        if (mode == SO_MODE_PE_2FILE) c.records_in[1] += (int64_t)ndealt;
#undef REC1
#undef REC2
    }

    free(b1.v); free(b2.v); free(cuts1); free(cuts2);
    for (int s = 0; s < 3; s++) {
        if (out_len) out_len[s] = ob[s].len;
        if (ob[s].overflow && !rc) rc = -1;
    }
    if (ctr) *ctr = c;
    if (err) *err = e;
    return rc;
}
