// cli.cpp -- `plass_b200_cli <command> <args>`: the GPU commands (the four hot-path steps + findassemblystart, cyclecheck) behind the reference's own
// command-line contract, so that `$MMSEQS kmermatcher ...` lines of data/assemble.sh / nuclassemble.sh can
// be pointed at this binary (see INTEGRATION.md for the dispatcher and for the Command-table stub).
#include "commands.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, const char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <kmermatcher|rescorediagonal|assembleresults|nuclassembleresults|findassemblystart|cyclecheck|extractorfs|translatenucs|assembleiteration|dbdiff> <dbs...> [flags]\n", argv[0]);
        return EXIT_FAILURE;
    }
    const char *cmd = argv[1];
    if (!strcmp(cmd, "kmermatcher")) return kmermatcher(argc - 2, argv + 2);
    if (!strcmp(cmd, "rescorediagonal")) return rescorediagonal(argc - 2, argv + 2);
    if (!strcmp(cmd, "assembleresults")) return assembleresults(argc - 2, argv + 2);
    if (!strcmp(cmd, "nuclassembleresults")) return nuclassembleresults(argc - 2, argv + 2);
    if (!strcmp(cmd, "findassemblystart")) return findassemblystart(argc - 2, argv + 2);
    if (!strcmp(cmd, "cyclecheck")) return cyclecheck(argc - 2, argv + 2);
    if (!strcmp(cmd, "extractorfs")) return extractorfs(argc - 2, argv + 2);
    if (!strcmp(cmd, "translatenucs")) return translatenucs(argc - 2, argv + 2);
    if (!strcmp(cmd, "assembleiteration")) return assembleiteration(argc - 2, argv + 2);
    if (!strcmp(cmd, "dbdiff")) return dbdiff(argc - 2, argv + 2);
    if (!strcmp(cmd, "iotest")) return iotest(argc - 2, argv + 2);
    fprintf(stderr, "%s: not one of the GPU hot-path commands\n", cmd);
    return EXIT_FAILURE;
}
