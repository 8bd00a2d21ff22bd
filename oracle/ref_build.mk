# Builds the UNMODIFIED reference (soedinglab/plass + vendored MMseqs2) from the sources where they
# lie under $(REF) into oracle/_ref/ -- test infrastructure only (the parity checker and the CPU
# baseline of bench.py).  This is our own recipe: the reference's CMake build system is not run.
# Nothing is copied into the repo; outputs (objects, generated resource headers, binaries) go to
# oracle/_ref/ only, which is git-ignored but travels to the GPU box.
#
#   make -f oracle/ref_build.mk -j8            # -> oracle/_ref/bin/{plass,penguin}
#
# SIMD level: -mavx2 (the reference's HAVE_AVX2 build, lib/mmseqs/CMakeLists.txt:52-58) so the
# binary runs on any AVX2 host, not only on the build container's CPU.
REF   ?= /root/reference
OUT   ?= $(dir $(lastword $(MAKEFILE_LIST)))_ref
OBJ   := $(OUT)/obj
GEN   := $(OUT)/generated
MM    := $(REF)/lib/mmseqs
CXX   := /usr/bin/g++
CC    := /usr/bin/gcc
PY    ?= python3

ARCH  := -mavx2 -mcx16
DEFS  := -DOPENMP=1 -DHAVE_ZLIB=1 -DENABLE_IPS4O=1 -DHAVE_POSIX_FADVISE=1 -DHAVE_POSIX_MADVISE=1 -DGIT_SHA1=ac83d8f-oracle
INCS  := -I$(GEN) -I$(MM)/src -I$(MM)/src/alignment -I$(MM)/src/clustering -I$(MM)/src/commons \
         -I$(MM)/src/multihit -I$(MM)/src/prefiltering -I$(MM)/src/linclust -I$(MM)/src/taxonomy \
         -I$(MM)/src/util -I$(MM)/lib -I$(MM)/lib/simd -I$(MM)/lib/simde -I$(MM)/lib/gzstream \
         -I$(MM)/lib/alp -I$(MM)/lib/cacode -I$(MM)/lib/ksw2 -I$(MM)/lib/xxhash -I$(MM)/lib/ips4o \
         -I$(MM)/lib/zstd/lib -I$(MM)/lib/zstd/lib/common -I$(MM)/lib/tinyexpr -I$(MM)/lib/microtar \
         -I$(REF)/lib -I$(REF)/src/commons -I$(REF)/src
CXXFLAGS := -O3 -std=c++1y -fsigned-char $(ARCH) -fopenmp -include cstdint -w $(DEFS) $(INCS)
CFLAGS   := -O3 -fsigned-char $(ARCH) -w -DZSTD_STATIC_LINKING_ONLY $(INCS)

FW_DIRS := alignment clustering commons linclust multihit prefiltering taxonomy util workflow
FW_SRC  := $(foreach d,$(FW_DIRS),$(wildcard $(MM)/src/$(d)/*.cpp)) $(MM)/src/MMseqsBase.cpp
FW_SRC  := $(filter-out %/multihit/resultsbyset.cpp,$(FW_SRC))
LIB_CXX := $(wildcard $(MM)/lib/alp/*.cpp) $(MM)/lib/ksw2/ksw2_extz2_sse.cpp $(wildcard $(MM)/lib/cacode/*.cpp) \
           $(wildcard $(REF)/lib/flash/*.cpp) $(wildcard $(REF)/lib/kerasify/*.cpp)
LIB_C   := $(MM)/lib/tinyexpr/tinyexpr.c $(MM)/lib/microtar/microtar.c \
           $(wildcard $(MM)/lib/zstd/lib/common/*.c) $(wildcard $(MM)/lib/zstd/lib/compress/*.c) \
           $(wildcard $(MM)/lib/zstd/lib/decompress/*.c)
COMMON_SRC := $(REF)/src/commons/LocalParameters.cpp $(REF)/src/util/createhdb.cpp $(REF)/src/version/Version.cpp \
              $(REF)/src/assembler/mergereads.cpp
PLASS_SRC   := $(REF)/src/plass.cpp $(REF)/src/workflow/Assembler.cpp $(REF)/src/assembler/assembleresult.cpp \
               $(REF)/src/assembler/findassemblystart.cpp $(REF)/src/assembler/filternoncoding.cpp
PENGUIN_SRC := $(REF)/src/penguin.cpp $(REF)/src/workflow/Nuclassembler.cpp $(REF)/src/workflow/GuidedNuclassembler.cpp \
               $(REF)/src/assembler/nuclassembleresult.cpp $(REF)/src/assembler/guidedassembleresult.cpp \
               $(REF)/src/assembler/cyclecheck.cpp

o = $(patsubst $(REF)/%,$(OBJ)/%.o,$(1))
FW_OBJ      := $(call o,$(FW_SRC) $(LIB_CXX) $(LIB_C))
COMMON_OBJ  := $(call o,$(COMMON_SRC))
PLASS_OBJ   := $(call o,$(PLASS_SRC))
PENGUIN_OBJ := $(call o,$(PENGUIN_SRC))

all: $(OUT)/bin/plass $(OUT)/bin/penguin

$(GEN)/.stamp: $(dir $(lastword $(MAKEFILE_LIST)))gen_resources.py
	mkdir -p $(GEN)
	$(PY) $< $(REF) $(GEN)
	touch $@

$(OBJ)/%.cpp.o: $(REF)/%.cpp $(GEN)/.stamp
	@mkdir -p $(dir $@)
	@echo CXX $< && $(CXX) $(CXXFLAGS) -c $< -o $@

$(OBJ)/%.c.o: $(REF)/%.c $(GEN)/.stamp
	@mkdir -p $(dir $@)
	@echo CC $< && $(CC) $(CFLAGS) -c $< -o $@

$(OUT)/libmmseqs-framework.a: $(FW_OBJ)
	@rm -f $@ && ar rcs $@ $^

$(OUT)/bin/plass: $(PLASS_OBJ) $(COMMON_OBJ) $(OUT)/libmmseqs-framework.a
	@mkdir -p $(dir $@)
	$(CXX) -fopenmp -o $@ $(PLASS_OBJ) $(COMMON_OBJ) $(OUT)/libmmseqs-framework.a -lz -latomic -lpthread

$(OUT)/bin/penguin: $(PENGUIN_OBJ) $(COMMON_OBJ) $(OUT)/libmmseqs-framework.a
	@mkdir -p $(dir $@)
	$(CXX) -fopenmp -o $@ $(PENGUIN_OBJ) $(COMMON_OBJ) $(OUT)/libmmseqs-framework.a -lz -latomic -lpthread

# The compiled reference-side binding: the reference's plass tool + three GPU Command rows.
# Needs plass_b200/libplassgpu.so (python -m plass_b200.build); still test infrastructure, lives beside the reference binaries.
REPO := $(abspath $(dir $(lastword $(MAKEFILE_LIST)))..)
SHIM_SRC := $(REPO)/oracle/shim/gpu_commands.cpp
$(OBJ)/shim/gpu_commands.o: $(SHIM_SRC) $(REPO)/include/plassgpu.h $(GEN)/.stamp
	@mkdir -p $(dir $@)
	@echo CXX $< && $(CXX) $(CXXFLAGS) -DPLASS_TOOL_CPP='"$(REF)/src/plass.cpp"' -I$(REPO)/include -c $< -o $@

shim: $(OUT)/bin/plass_gpu_shim
$(OUT)/bin/plass_gpu_shim: $(OBJ)/shim/gpu_commands.o $(filter-out %/src/plass.cpp.o,$(PLASS_OBJ)) $(COMMON_OBJ) $(OUT)/libmmseqs-framework.a $(REPO)/plass_b200/libplassgpu.so
	@mkdir -p $(dir $@)
	$(CXX) -fopenmp -o $@ $(OBJ)/shim/gpu_commands.o $(filter-out %/src/plass.cpp.o,$(PLASS_OBJ)) $(COMMON_OBJ) $(OUT)/libmmseqs-framework.a \
	    -L$(REPO)/plass_b200 -lplassgpu -Wl,-rpath,'$$ORIGIN/../../../plass_b200' -L/usr/local/cuda/lib64 -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -lz -latomic -lpthread

clean:
	rm -rf $(OBJ) $(GEN) $(OUT)/bin $(OUT)/libmmseqs-framework.a
.PHONY: all clean shim
