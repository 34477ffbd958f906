# Builds libspgpu.so (CUDA kernels + C ABI) for sm_100a, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas --expt-relaxed-constexpr
PKG := spartan_parallel_b200
SRC := $(wildcard $(PKG)/csrc/*.cu)
OBJ := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRC))
HDR := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/spgpu.h
LIB := $(PKG)/libspgpu.so

HOSTLIB := $(PKG)/libspghost.so
HOSTSRC := $(wildcard $(PKG)/host/*.cpp) $(wildcard $(PKG)/host/*.hpp)

all: $(LIB) $(HOSTLIB)

# C++ mirror of the reference's transcript-side host code, above the C ABI
$(HOSTLIB): $(HOSTSRC) $(PKG)/csrc/ed25519.cuh $(PKG)/csrc/host_fq.h include/spgpu.h $(LIB)
	g++ -O3 -std=c++17 -fPIC -shared -Wall -Wno-unknown-pragmas -o $@ $(PKG)/host/capi.cpp -L$(PKG) -lspgpu -Wl,-rpath,'$$ORIGIN'

build/%.o: $(PKG)/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

oracle:
	$(MAKE) -C oracle

tools: build/imad_peak build/round_latency

build/round_latency: tools/round_latency.cu $(PKG)/csrc/fq.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ $<

build/imad_peak: tools/imad_peak.cu $(PKG)/csrc/fq.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ $<

clean:
	rm -rf build $(LIB) $(HOSTLIB)

.PHONY: all clean oracle tools
