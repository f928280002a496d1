# Builds libbndm_b200.so (sm_100a) in-tree.  `make` / `python -c "import __graft_entry__ as g; g.build()"`.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
SRC       := $(wildcard bndm_b200/csrc/*.cu)
OBJ       := $(patsubst bndm_b200/csrc/%.cu,build/%.o,$(SRC))
LIB       := bndm_b200/lib/libbndm_b200.so

all: $(LIB)

build/%.o: bndm_b200/csrc/%.cu $(wildcard bndm_b200/csrc/*.cuh) include/bndm_b200.h
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p bndm_b200/lib
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

# stand-alone measurement probes (not part of the library): `make probes` here, then run build/<probe> under gpurun
probes: build/stream_probe build/get_noise_probe build/pipe_probe
build/get_noise_probe: tools/probes/get_noise_probe.cu $(LIB) include/bndm_b200.h
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -o $@ $< -Lbndm_b200/lib -lbndm_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../bndm_b200/lib'
build/stream_probe: tools/probes/stream_probe.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ $<

build/pipe_probe: tools/probes/pipe_probe.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ $<

clean:
	rm -rf build $(LIB)
.PHONY: all clean probes
