/*
 * sickle_oracle_main.c -- tiny command-line front end for the CPU oracle.
 * TEST INFRASTRUCTURE ONLY (see sickle_oracle.h).  Usage mirrors the reference CLI subset
 * the tests need:
 *   sickle_oracle se -f in.fq -t sanger -o out.fq [-q N] [-l N] [-x] [-n] [-a N] [-b MiB]
 *   sickle_oracle pe -f a.fq -r b.fq -t T -o o1 -p o2 -s singles [...]
 *   sickle_oracle pe -c inter.fq -t T -m out [-s singles] | -M out
 * Plain (uncompressed) input only.
 */
#include "sickle_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static char *slurp(const char *path, size_t *n) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)malloc((size_t)sz + 1);
    if (sz > 0 && fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "short read %s\n", path); exit(2); }
    fclose(f);
    *n = (size_t)sz;
    return buf;
}

static void spill(const char *path, const char *buf, size_t n) {
    if (!path) return;
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path); exit(2); }
    if (n) fwrite(buf, 1, n, f);
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 2 || (strcmp(argv[1], "se") && strcmp(argv[1], "pe"))) {
        fprintf(stderr, "usage: %s se|pe [options]\n", argv[0]);
        return 1;
    }
    int pe = !strcmp(argv[1], "pe");
    const char *f = NULL, *r = NULL, *c = NULL, *o = NULL, *p2 = NULL, *m = NULL, *M = NULL, *s = NULL;
    so_params p = {-1, 20, 20, 0, 0};
    int threads = 1;
    long b_mib = 512;
    for (int i = 2; i < argc; i++) {
        const char *a = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if (!strcmp(a, "-x")) p.no_fiveprime = 1;
        else if (!strcmp(a, "-n")) p.trunc_n = 1;
        else if (!v) { fprintf(stderr, "missing value for %s\n", a); return 1; }
        else if (!strcmp(a, "-f")) { f = v; i++; }
        else if (!strcmp(a, "-r")) { r = v; i++; }
        else if (!strcmp(a, "-c")) { c = v; i++; }
        else if (!strcmp(a, "-o")) { o = v; i++; }
        else if (!strcmp(a, "-p")) { p2 = v; i++; }
        else if (!strcmp(a, "-m")) { m = v; i++; }
        else if (!strcmp(a, "-M")) { M = v; i++; }
        else if (!strcmp(a, "-s")) { s = v; i++; }
        else if (!strcmp(a, "-q")) { p.qual_threshold = atoi(v); i++; }
        else if (!strcmp(a, "-l")) { p.length_threshold = atoi(v); i++; }
        else if (!strcmp(a, "-a")) { threads = atoi(v); i++; }
        else if (!strcmp(a, "-b")) { b_mib = atol(v); i++; }
        else if (!strcmp(a, "-t")) {
            if (!strcmp(v, "sanger")) p.qualtype = SO_SANGER;
            else if (!strcmp(v, "solexa")) p.qualtype = SO_SOLEXA;
            else if (!strcmp(v, "illumina")) p.qualtype = SO_ILLUMINA;
            i++;
        } else { fprintf(stderr, "unknown option %s\n", a); return 1; }
    }
    if (p.qualtype < 0) { fprintf(stderr, "need -t\n"); return 1; }

    int mode;
    size_t n1 = 0, n2 = 0;
    char *in1 = NULL, *in2 = NULL;
    if (!pe) { mode = SO_MODE_SE; in1 = slurp(f, &n1); }
    else if (c) { mode = M ? SO_MODE_PE_INTER_M : SO_MODE_PE_INTER; in1 = slurp(c, &n1); }
    else { mode = SO_MODE_PE_2FILE; in1 = slurp(f, &n1); in2 = slurp(r, &n2); }

    char *out[3];
    size_t cap[3], len[3];
    for (int k = 0; k < 3; k++) { cap[k] = n1 + n2 + 16; out[k] = (char *)malloc(cap[k]); }
    so_counters ctr;
    so_error err;
    int64_t bl = so_recommended_batch_len((int64_t)n1, b_mib, pe);
    int rc = so_run(mode, &p, threads, bl, s != NULL, in1, n1, in2, n2, out, cap, len, &ctr, &err);

    if (mode == SO_MODE_SE) spill(o, out[0], len[0]);
    else if (mode == SO_MODE_PE_2FILE) { spill(o, out[0], len[0]); spill(p2, out[1], len[1]); spill(s, out[2], len[2]); }
    else if (mode == SO_MODE_PE_INTER) { spill(m, out[0], len[0]); spill(s, out[2], len[2]); }
    else spill(M, out[0], len[0]);

    if (rc) {
        fprintf(stderr, "oracle: error kind %d at record %lld (file %d) position %d byte %d\n", err.kind,
                (long long)err.record, err.file, err.position, err.byte);
        return 1;
    }
    if (!pe) printf("records %lld kept %lld discarded %lld\n", (long long)(ctr.kept + ctr.discard),
                    (long long)ctr.kept, (long long)ctr.discard);
    else printf("kept_p %lld kept_s1 %lld kept_s2 %lld discard_p %lld discard_s1 %lld discard_s2 %lld\n",
                (long long)ctr.kept_p, (long long)ctr.kept_s1, (long long)ctr.kept_s2, (long long)ctr.discard_p,
                (long long)ctr.discard_s1, (long long)ctr.discard_s2);
    return 0;
}
