// commands.h -- the four drop-in commands.  Same signature as the reference's command functions
// (lib/mmseqs/src/commons/Command.h:91-102: int fn(int argc, const char **argv, const Command &)),
// minus the Command descriptor: argv excludes program and command name (Application.cpp:203).
#pragma once
int kmermatcher(int argc, const char **argv);            // replaces linclust/kmermatcher.cpp:780
int rescorediagonal(int argc, const char **argv);        // replaces alignment/rescorediagonal.cpp:381
int assembleresults(int argc, const char **argv);        // replaces src/assembler/assembleresult.cpp:358
int nuclassembleresults(int argc, const char **argv);    // replaces src/assembler/nuclassembleresult.cpp:400
int findassemblystart(int argc, const char **argv);      // replaces src/assembler/findassemblystart.cpp:35
int cyclecheck(int argc, const char **argv);             // replaces src/assembler/cyclecheck.cpp:31
int extractorfs(int argc, const char **argv);            // replaces lib/mmseqs/src/util/extractorfs.cpp:20
int translatenucs(int argc, const char **argv);          // replaces lib/mmseqs/src/util/translatenucs.cpp:14
int assembleiteration(int argc, const char **argv);      // the three hot-path steps of one iteration fused in one process
int dbdiff(int argc, const char **argv);                 // logical key -> entry comparison of two DBs (test / bench tool)
int iotest(int argc, const char **argv);                 // text layer round trip without a GPU (CPU tests, host-side timing)
