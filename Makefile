# Top-level build: product library (CUDA, sm_100a) + test oracle (gcc).  `make -j8`.
#   fastlanes_b200/lib/libfastlanes_b200.so   — the C-ABI product (include/fastlanes_b200.h)
#   oracle/_build/libfl_oracle.so             — CPU oracle (test infrastructure; never linked by the product)
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS ?= -std=c++17 -O3 -lineinfo $(ARCH) -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
           --expt-relaxed-constexpr
SRC := fastlanes_b200/csrc
OBJ := build/obj
LIB := fastlanes_b200/lib/libfastlanes_b200.so
TYPES := 64 32 16 8
PARTS := 0 1 2 3
HDRS := $(SRC)/fl_device.cuh $(SRC)/fl_kernels.cuh $(SRC)/fl_scan.cuh $(SRC)/fl_scan_bits.h $(SRC)/fl_internal.h include/fastlanes_b200.h

CODEC_OBJS := $(foreach t,$(TYPES),$(foreach p,$(PARTS),$(OBJ)/codec_u$(t)_p$(p).o))
OBJS := $(CODEC_OBJS) $(OBJ)/fl_misc.o $(OBJ)/fl_api.o

all: lib oracle build/test_traits
lib: $(LIB)
oracle:
	$(MAKE) -C oracle

$(OBJ):
	mkdir -p $(OBJ) fastlanes_b200/lib

define CODEC_RULE
$(OBJ)/codec_u$(1)_p$(2).o: $(SRC)/fl_codec_inst.cu $(HDRS) | $(OBJ)
	$(NVCC) $(NVFLAGS) -DFLB_TBITS=$(1) -DFLB_PART=$(2) -c $$< -o $$@
endef
$(foreach t,$(TYPES),$(foreach p,$(PARTS),$(eval $(call CODEC_RULE,$(t),$(p)))))

$(OBJ)/fl_misc.o: $(SRC)/fl_misc.cu $(HDRS) | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(OBJ)/fl_api.o: $(SRC)/fl_api.cu $(HDRS) | $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -cudart static

tools/kbench: tools/kbench.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -I$(SRC) -o $@ $< -cudart static

clean:
	rm -rf build fastlanes_b200/lib tools/kbench
	$(MAKE) -C oracle clean

.PHONY: all lib oracle clean

# kernel micro-benchmark variants: make kbench-variants
KB := build/kbench
$(KB):
	mkdir -p $(KB)
$(KB)/kb_base: tools/kbench.cu $(HDRS) | $(KB)
	$(NVCC) $(NVFLAGS) -I$(SRC) -o $@ $< -cudart static
$(KB)/kb_%: tools/kbench.cu $(HDRS) | $(KB)
	$(NVCC) $(NVFLAGS) -I$(SRC) $(KBFLAGS_$*) -o $@ $< -cudart static
KBFLAGS_st1 := -DFLB_ST_MODE=1
KBFLAGS_st2 := -DFLB_ST_MODE=2
KBFLAGS_ld1 := -DFLB_LD_MODE=1
KBFLAGS_ld3 := -DFLB_LD_MODE=3
KBFLAGS_pf4 := -DFLB_PREFETCH=4
KBFLAGS_pf16 := -DFLB_PREFETCH=16
KBFLAGS_t128 := -DFLB_THREADS=128
KBFLAGS_t512 := -DFLB_THREADS=512
kbench-variants: $(KB)/kb_base $(KB)/kb_st1 $(KB)/kb_st2 $(KB)/kb_ld1 $(KB)/kb_ld3 $(KB)/kb_pf4 $(KB)/kb_pf16 $(KB)/kb_t128 $(KB)/kb_t512
KBFLAGS_u32 := -DKB_ONLY_U32

# C++ trait-mirror test binary (run on the GPU box by tests/test_gpu_cpp_traits.py)
build/test_traits: tests/cpp/test_traits.cpp include/fastlanes_b200.hpp include/fastlanes_b200.h $(LIB)
	g++ -std=c++17 -O1 -Iinclude -o $@ $< -Lfastlanes_b200/lib -lfastlanes_b200 -Wl,-rpath,'$$ORIGIN/../fastlanes_b200/lib'

# CPU check of the scan kernels' bit arithmetic (warp emulated in lockstep; no CUDA): tests/test_scan_bits.py
build/test_scan_bits: tests/cpp/test_scan_bits.cpp $(SRC)/fl_scan_bits.h
	mkdir -p build
	g++ -std=c++17 -O1 -Wall -I$(SRC) -o $@ $<

# latency of the single-block drop-in call from compiled host code (tools/latbench.cpp)
build/latbench: tools/latbench.cpp include/fastlanes_b200.h $(LIB)
	g++ -std=c++17 -O2 -Iinclude -o $@ $< -Lfastlanes_b200/lib -lfastlanes_b200 -Wl,-rpath,'$$ORIGIN/../fastlanes_b200/lib'

# mid-size host calls (the reference's 1024-block throughput bench shape) from compiled host code (tools/midbench.cpp)
build/midbench: tools/midbench.cpp include/fastlanes_b200.h $(LIB)
	g++ -std=c++17 -O2 -Iinclude -o $@ $< -Lfastlanes_b200/lib -lfastlanes_b200 -Wl,-rpath,'$$ORIGIN/../fastlanes_b200/lib'
