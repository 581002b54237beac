# Top-level build: libpcf.so (CUDA, sm_100a), the five drop-in front ends under bin/, the oracle.
NVCC      ?= /usr/local/cuda/bin/nvcc
HOSTCXX   := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -ccbin $(HOSTCXX) -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off
CSRC      := parcompfin_b200/csrc
SRCS      := $(CSRC)/pcf_api.cu $(CSRC)/mc_kernels.cu $(CSRC)/amer_kernels.cu $(CSRC)/binom_kernels.cu $(CSRC)/tree_kernels.cu $(CSRC)/peaks.cu
OBJS      := $(SRCS:.cu=.o) $(CSRC)/fastmath_tables.o
HDRS      := $(wildcard $(CSRC)/*.cuh) include/pcf.h
LIB       := parcompfin_b200/libpcf.so
HOST      := parcompfin_b200/host
BINS      := bin/mc_eur bin/mc_eur_multi bin/mc_asia bin/mc_amer bin/binom_embar bin/binom_vanilla_eur bin/binom_vanilla_amer

.PHONY: all lib bins oracle clean
all: lib bins oracle

lib: $(LIB)
$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(@:.o=.ptxas.log) || { cat $(@:.o=.ptxas.log); exit 1; }
$(CSRC)/fastmath_tables.o: $(CSRC)/fastmath_tables.cpp $(CSRC)/fastmath.cuh
	$(HOSTCXX) -std=c++17 -O2 -fPIC -fvisibility=hidden -ffp-contract=off -c $< -o $@
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(HOSTCXX) -o $@ $(OBJS) -ldl -lpthread

bins: $(BINS)
bin/%: $(HOST)/%.cpp $(HOST)/frontend.h include/pcf.h $(LIB)
	@mkdir -p bin
	$(HOSTCXX) -std=c++17 -O2 -Iinclude -I$(HOST) $< -o $@ -Lparcompfin_b200 -lpcf -Wl,-rpath,'$$ORIGIN/../parcompfin_b200'

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(OBJS) $(CSRC)/*.ptxas.log $(LIB) $(BINS)
