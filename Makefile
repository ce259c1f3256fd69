# Makefile — same interface as the reference's (Makefile:10-23,39-62):
#   make KERNEL=<name> [NVCC_ARCH=100a] [MAXRREGCOUNT=n]   ->  bin/profile_<name>
# plus
#   make lib        -> quantizedmha_b200/lib/libqmha.so   (the C-ABI shared library)
#   make oracle     -> oracle/libqmha_oracle.so (+ oracle/_ref when /root/reference exists)
# KERNEL accepts the reference's names (fa_tc_int8_b, fa_tc_v2a, fa, unfused, ...) and the
# native ones (fa_b200_int8, fa_b200_f16); it only selects the DEFAULT variant of solve() —
# every variant is compiled into every binary.
KERNEL    ?= fa_tc_int8_b
NVCC_ARCH ?= 100a
NVCC      ?= nvcc
# Blackwell tcgen05/TMA code needs the arch-specific target: an explicit gencode pair, never
# plain -arch=sm_100a (that also emits generic compute_100 PTX, which ptxas rejects).
GENCODE   := -gencode arch=compute_$(NVCC_ARCH),code=sm_$(NVCC_ARCH)
NVCCFLAGS := -O3 -std=c++17 -lineinfo $(GENCODE) -Xcompiler -fPIC --ptxas-options=-v
ifdef MAXRREGCOUNT
NVCCFLAGS += -maxrregcount=$(MAXRREGCOUNT)
endif

CSRC    := quantizedmha_b200/csrc
LIBDIR  := quantizedmha_b200/lib
OBJDIR  := build/obj
LIB     := $(LIBDIR)/libqmha.so
KOBJS   := $(OBJDIR)/attn_fwd.o $(OBJDIR)/prepare.o
HDRS    := $(wildcard $(CSRC)/*.cuh) include/qmha.h

all: lib driver

lib: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; false)

# api_<KERNEL>.o carries the default variant of solve(), so there is one per KERNEL.
$(OBJDIR)/api_%.o: $(CSRC)/api.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -DQMHA_DEFAULT_KERNEL='"$*"' -c $< -o $@

$(LIB): $(KOBJS) $(OBJDIR)/api_fa_tc_int8_b.o
	@mkdir -p $(LIBDIR)
	$(NVCC) -shared $(GENCODE) -o $@ $^

driver: bin/profile_$(KERNEL)

bin/profile_$(KERNEL): drivers/main.cu inputs/data.cu utils/verify.cu $(KOBJS) $(OBJDIR)/api_$(KERNEL).o
	@mkdir -p bin
	$(NVCC) $(NVCCFLAGS) -DQMHA_DEFAULT_KERNEL='"$(KERNEL)"' -o $@ drivers/main.cu inputs/data.cu utils/verify.cu $(KOBJS) $(OBJDIR)/api_$(KERNEL).o

oracle:
	$(MAKE) -C oracle
	@if [ -d /root/reference ]; then $(MAKE) -C oracle ref; fi

# compute-sanitizer on a tiny shape (README.md:158-194 of the reference documents the same tools)
sanitize: bin/profile_$(KERNEL)
	compute-sanitizer --tool memcheck bin/profile_$(KERNEL) --N=256 --d_model=128 --h=2 --warmup=0 --runs=1

clean:
	rm -rf build bin $(LIBDIR)/libqmha.so

.PHONY: all lib driver oracle sanitize clean
