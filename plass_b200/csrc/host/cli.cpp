// cli.cpp -- `plass_b200_cli <command> <args>`: the GPU commands (the four hot-path steps + findassemblystart, cyclecheck) behind the reference's own
// command-line contract, so that `$MMSEQS kmermatcher ...` lines of data/assemble.sh / nuclassemble.sh can
// be pointed at this binary (see INTEGRATION.md for the dispatcher and for the Command-table stub).
#include "commands.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

int main(int argc, const char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <kmermatcher|rescorediagonal|assembleresults|nuclassembleresults|findassemblystart|cyclecheck|extractorfs|translatenucs|assembleiteration|dbdiff> <dbs...> [flags]\n", argv[0]);
        return EXIT_FAILURE;
    }
    const char *cmd = argv[1];
    int rc = -1;
    if (!strcmp(cmd, "kmermatcher")) rc = kmermatcher(argc - 2, argv + 2);
    else if (!strcmp(cmd, "rescorediagonal")) rc = rescorediagonal(argc - 2, argv + 2);
    else if (!strcmp(cmd, "assembleresults")) rc = assembleresults(argc - 2, argv + 2);
    else if (!strcmp(cmd, "nuclassembleresults")) rc = nuclassembleresults(argc - 2, argv + 2);
    else if (!strcmp(cmd, "findassemblystart")) rc = findassemblystart(argc - 2, argv + 2);
    else if (!strcmp(cmd, "cyclecheck")) rc = cyclecheck(argc - 2, argv + 2);
    else if (!strcmp(cmd, "extractorfs")) rc = extractorfs(argc - 2, argv + 2);
    else if (!strcmp(cmd, "translatenucs")) rc = translatenucs(argc - 2, argv + 2);
    else if (!strcmp(cmd, "assembleiteration")) rc = assembleiteration(argc - 2, argv + 2);
    else if (!strcmp(cmd, "dbdiff")) return dbdiff(argc - 2, argv + 2);
    else if (!strcmp(cmd, "iotest")) return iotest(argc - 2, argv + 2);
    if (rc >= 0) {
        // the DBs are written and closed: skip the teardown of the CUDA context (it costs more than some commands)
        fflush(stdout);
        fflush(stderr);
        _exit(rc);
    }
    fprintf(stderr, "%s: not one of the GPU hot-path commands\n", cmd);
    return EXIT_FAILURE;
}
