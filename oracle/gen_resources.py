#!/usr/bin/env python3
"""Turn the reference's data files (substitution matrices, workflow shell scripts, resource
libraries, keras models) into the `<name>.h` byte-array headers its sources #include.

The reference does this at configure time with `xxd -i` + sed
(lib/mmseqs/cmake/MMseqsResourceCompiler.cmake:35-50); this is our own equivalent so that
oracle/ref_build.mk does not have to run the reference's CMake build system.  Output goes to
oracle/_ref/generated/ only (never into the repo history).

usage: gen_resources.py <reference root> <output dir>
"""
import os
import re
import sys


def emit(src, out_dir):
    name = os.path.basename(src)
    sym = re.sub(r"[^0-9A-Za-z]", "_", name)
    if sym[0].isdigit():
        sym = "__" + sym
    data = open(src, "rb").read()
    lines = []
    for i in range(0, len(data), 12):
        lines.append("  " + ", ".join("0x%02x" % b for b in data[i:i + 12]))
    body = ",\n".join(lines)
    with open(os.path.join(out_dir, name + ".h"), "w") as f:
        f.write("static const unsigned char %s[] = {\n%s\n};\n" % (sym, body))
        f.write("unsigned int %s_len = %d;\n" % (sym, len(data)))


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    mm = os.path.join(ref, "lib", "mmseqs", "data")
    files = []
    for n in ("VTML80.out", "VTML40.out", "nucleotide.out", "blosum62.out", "PAM30.out"):
        files.append(os.path.join(mm, n))
    for sub in ("resources", "workflow"):
        d = os.path.join(mm, sub)
        for n in sorted(os.listdir(d)):
            if n != "CMakeLists.txt":
                files.append(os.path.join(d, n))
    d = os.path.join(ref, "data")
    for n in sorted(os.listdir(d)):
        if n != "CMakeLists.txt":
            files.append(os.path.join(d, n))
    for f in files:
        emit(f, out)
    print("generated %d resource headers in %s" % (len(files), out))


if __name__ == "__main__":
    main()
