// sickle_main.cpp -- `sickle <se|pe|--help|--version>` dispatcher (reference src/sickle.cpp:40-86).
// Same commands, usage text and exit codes; the trimmers run on the GPU through include/sickle_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <getopt.h>
#include <unistd.h>

#include "trimmer.h"

static void main_usage(int status) {
    fprintf(stdout, "\nUsage: %s <command> [options]\n\
\n\
Command:\n\
pe\tpaired-end sequence trimming\n\
se\tsingle-end sequence trimming\n\
\n\
--help, display this help and exit\n\
--version, output version information and exit\n\n", PROGRAM_NAME);
    exit(status);
}

static int run_command(int argc, char *argv[]) {
    if (strcmp(argv[1], "pe") == 0) {
        Trim_Paired trimmer;
        const int rc = trimmer.parse_args(argc, argv);
        return rc != 0 ? rc : trimmer.trim_main();
    }
    Trim_Single trimmer;
    const int rc = trimmer.parse_args(argc, argv);
    return rc != 0 ? rc : trimmer.trim_main();
}

// `sickle batch` (not in the reference): one `se ...` / `pe ...` command per line on stdin, run one
// after the other in this process, so that CUDA start-up and the pinned buffers are paid once per
// process instead of once per file.  After every command a line `##rc <exit code>` goes to stdout.
// Arguments are split at blanks; double quotes keep blanks.  A usage error ends the whole batch, as
// it would end `sickle` itself.
static int run_batch() {
    host::batch_mode = true;
    int worst = EXIT_SUCCESS;
    char *line = nullptr;
    size_t cap = 0;
    while (getline(&line, &cap, stdin) > 0) {
        std::vector<std::string> tok;
        std::string cur;
        bool quoted = false, any = false;
        for (const char *p = line; *p && *p != '\n'; ++p) {
            if (*p == '"') { quoted = !quoted; any = true; }
            else if ((*p == ' ' || *p == '\t') && !quoted) { if (any) tok.push_back(cur); cur.clear(); any = false; }
            else { cur += *p; any = true; }
        }
        if (any) tok.push_back(cur);
        if (tok.empty()) continue;
        if (tok[0] != "se" && tok[0] != "pe") { fprintf(stderr, "****Error: batch lines start with se or pe\n\n"); worst = EXIT_FAILURE; continue; }
        std::vector<char *> av;
        std::string prog = PROGRAM_NAME;
        av.push_back(&prog[0]);
        for (auto &t : tok) av.push_back(&t[0]);
        av.push_back(nullptr);
        optind = 0;   // glibc: restart getopt_long from scratch
        const int rc = run_command((int)av.size() - 1, av.data());
        fprintf(stdout, "##rc %d\n", rc);
        fflush(stdout);
        if (rc != EXIT_SUCCESS) worst = rc;
    }
    fflush(NULL);
    _exit(worst);
}

int main(int argc, char *argv[]) {
    if (argc >= 2 && strcmp(argv[1], "batch") == 0) return run_batch();
    if (argc < 2 || (strcmp(argv[1], "pe") != 0 && strcmp(argv[1], "se") != 0 && strcmp(argv[1], "--version") != 0 &&
                     strcmp(argv[1], "--help") != 0))
        main_usage(EXIT_FAILURE);
    if (strcmp(argv[1], "--version") == 0) {
        fprintf(stdout, "%s version %0.2f\nCopyright (c) 2011 The Regents of University of California, \
		Davis Campus.\n%s is free software and comes with ABSOLUTELY NO WARRANTY.\nDistributed under the\
		 MIT License.\n\nWritten by %s\n", PROGRAM_NAME, (double)SICKLE_VERSION, PROGRAM_NAME,
                "Nikhil Joshi, UC Davis Bioinformatics Core\n");
        return EXIT_SUCCESS;
    }
    if (strcmp(argv[1], "--help") == 0) main_usage(EXIT_SUCCESS);

    const int retval = run_command(argc, argv);
    // Every output file is closed by now.  Leave without the CUDA runtime's exit handlers: tearing the
    // context down call by call costs several hundred milliseconds that the kernel's cleanup does not.
    fflush(NULL);
    _exit(retval);
}
