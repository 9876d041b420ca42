// sickle_main.cpp -- `sickle <se|pe|--help|--version>` dispatcher (reference src/sickle.cpp:40-86).
// Same commands, usage text and exit codes; the trimmers run on the GPU through include/sickle_b200.h.
#include <cstdio>
#include <cstdlib>
#include <cstring>

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

int main(int argc, char *argv[]) {
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

    int retval;
    if (strcmp(argv[1], "pe") == 0) {
        Trim_Paired trimmer;
        retval = trimmer.parse_args(argc, argv);
        if (retval != 0) return retval;
        retval = trimmer.trim_main();
    } else {
        Trim_Single trimmer;
        retval = trimmer.parse_args(argc, argv);
        if (retval != 0) return retval;
        retval = trimmer.trim_main();
    }
    // Every output file is closed by now.  Leave without the CUDA runtime's exit handlers: tearing the
    // context down call by call costs several hundred milliseconds that the kernel's cleanup does not.
    fflush(NULL);
    _exit(retval);
}
