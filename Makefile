# Top-level build: the CUDA library (C ABI), the `sickle` CLI host, and the test oracle.
#   make lib     -> sickle_b200/libsickle_b200.so   (sm_100a only)
#   make cli     -> bin/sickle                       (drop-in `sickle se|pe` host over the C ABI)
#   make oracle  -> oracle/_build/*, oracle/_ref/* (the latter only where /root/reference exists)
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC      := sickle_b200/csrc
LIB       := sickle_b200/libsickle_b200.so
KERNELS   := $(wildcard $(CSRC)/*.cuh) include/sickle_b200.h

.PHONY: all lib cli oracle clean variant
all: lib cli oracle

lib: $(LIB)
$(LIB): $(CSRC)/capi.cu $(KERNELS)
	$(NVCC) $(NVFLAGS) -shared $(CSRC)/capi.cu -o $@ 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; exit 1)
	@grep -E "error|warning: v|spill" $(CSRC)/ptxas.log | grep -v "0 bytes spill" || true

# experimental build variants for same-GPU A/B runs (profiles/ab.sh): make variant V=SK_LANE_SPLIT4
#   -> sickle_b200/libsickle_b200_SK_LANE_SPLIT4.so   (select with SICKLE_B200_LIB=...)
variant:
	$(NVCC) $(NVFLAGS) -D$(V) -shared $(CSRC)/capi.cu -o sickle_b200/libsickle_b200_$(V).so 2> $(CSRC)/ptxas_$(V).log || (cat $(CSRC)/ptxas_$(V).log; exit 1)
	@grep -E "spill" $(CSRC)/ptxas_$(V).log | grep -v "0 bytes spill stores, 0 bytes spill loads" || true

cli: bin/sickle bin/io_tool
# the CLI's I/O stages alone (no CUDA): used by tests/test_host_io.py
bin/io_tool: host/io_tool.cpp host/io.cpp host/io.h host/ref_batcher.h host/unit_cutter.h
	@mkdir -p bin
	g++ -O2 -std=c++17 -Wall host/io_tool.cpp host/io.cpp -o $@ -lz -lpthread
bin/sickle: host/sickle_main.cpp host/trimmer.cpp host/io.cpp host/trimmer.h host/io.h host/ref_batcher.h host/unit_cutter.h include/sickle_b200.h $(LIB)
	@mkdir -p bin
	g++ -O2 -std=c++17 -Wall -Iinclude host/sickle_main.cpp host/trimmer.cpp host/io.cpp -o $@ \
	    -Lsickle_b200 -lsickle_b200 -lz -lpthread -Wl,-rpath,'$$ORIGIN/../sickle_b200'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -rf $(LIB) bin $(CSRC)/ptxas.log
	$(MAKE) -C oracle clean
