# Builds libspgpu.so (CUDA kernels + C ABI) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas --expt-relaxed-constexpr
PKG := spartan_parallel_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
HDR := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/spgpu.h
LIB := $(PKG)/libspgpu.so

all: $(LIB)

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

oracle:
	$(MAKE) -C oracle

tools: build/imad_peak

build/imad_peak: tools/imad_peak.cu $(PKG)/csrc/fq.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ $<

clean:
	rm -rf build $(LIB)

.PHONY: all clean oracle tools
