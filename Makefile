# Builds the product (libswb200.so, sm_100a only), the issue-rate microbenchmark and the checker.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
CSRC      := mini_parallel_b200/csrc
LIB       := mini_parallel_b200/libswb200.so

all: $(LIB) variants build/rustseq_mini build/issue_rate_bench build/cell_loop_bench build/inflate_bench build/gunzip_bench oracle

$(LIB): $(CSRC)/swb_kernels.cu $(CSRC)/swb_fastq_kernels.cu $(CSRC)/swb_traceback.cu $(CSRC)/swb_capi.cu $(CSRC)/rustseq_host.cpp $(CSRC)/swb_kernels.cuh $(CSRC)/swb_inflate.cuh $(CSRC)/host_gunzip.h $(CSRC)/host_pgunzip.h include/swb200.h include/rustseq_host.h
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/swb_kernels.cu $(CSRC)/swb_fastq_kernels.cu $(CSRC)/swb_traceback.cu $(CSRC)/swb_capi.cu $(CSRC)/rustseq_host.cpp -lz

# TEST-ONLY build with every short-read / long-pair kernel variant (-DSWB_ALL_VARIANTS): the parity tests load it beside
# the product library and cross-check the variants against the oracle.  Nothing in the product loads it.
VARLIB := tests/native/libswb200_variants.so
variants: $(VARLIB)
$(VARLIB): $(CSRC)/swb_kernels.cu $(CSRC)/swb_fastq_kernels.cu $(CSRC)/swb_traceback.cu $(CSRC)/swb_capi.cu $(CSRC)/rustseq_host.cpp $(CSRC)/swb_kernels.cuh $(CSRC)/swb_inflate.cuh $(CSRC)/host_gunzip.h $(CSRC)/host_pgunzip.h include/swb200.h include/rustseq_host.h
	mkdir -p tests/native
	$(NVCC) $(NVCCFLAGS) -DSWB_ALL_VARIANTS -shared -o $@ $(CSRC)/swb_kernels.cu $(CSRC)/swb_fastq_kernels.cu $(CSRC)/swb_traceback.cu $(CSRC)/swb_capi.cu $(CSRC)/rustseq_host.cpp -lz

# the reference's CLI (main.rs) over the library; finds libswb200.so next to the package via rpath
build/rustseq_mini: $(CSRC)/rustseq_mini_main.cpp $(LIB)
	mkdir -p build
	g++ -O2 -std=c++17 -o $@ $(CSRC)/rustseq_mini_main.cpp -Lmini_parallel_b200 -lswb200 -Wl,-rpath,'$$ORIGIN/../mini_parallel_b200'

build/issue_rate_bench: $(CSRC)/issue_rate_bench.cu
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

build/cell_loop_bench: $(CSRC)/cell_loop_bench.cu
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -o $@ $<

# the BGZF inflate kernel and the FASTQ index kernels alone (self-contained: compiles the kernels in)
build/inflate_bench: tools/inflate_bench.cu $(CSRC)/swb_fastq_kernels.cu $(CSRC)/swb_kernels.cuh $(CSRC)/swb_inflate.cuh include/swb200.h
	mkdir -p build
	$(NVCC) $(ARCH) -O3 -lineinfo -std=c++17 -o $@ tools/inflate_bench.cu $(CSRC)/swb_fastq_kernels.cu -lz

# host gzip readers alone (serial, zlib, several threads per file): no GPU, no nvcc
build/gunzip_bench: tools/gunzip_bench.cpp $(CSRC)/host_gunzip.h $(CSRC)/host_pgunzip.h
	mkdir -p build
	g++ -O2 -std=c++17 -pthread -o $@ tools/gunzip_bench.cpp -lz

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(VARLIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean variants
