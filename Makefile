# Builds the product (libswb200.so, sm_100a only), the issue-rate microbenchmark and the checker.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
CSRC      := mini_parallel_b200/csrc
LIB       := mini_parallel_b200/libswb200.so

all: $(LIB) build/issue_rate_bench oracle

$(LIB): $(CSRC)/swb_kernels.cu $(CSRC)/swb_capi.cu $(CSRC)/swb_kernels.cuh include/swb200.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/swb_kernels.cu $(CSRC)/swb_capi.cu

build/issue_rate_bench: $(CSRC)/issue_rate_bench.cu
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
