# Build of the sm_100a CUDA library, the C oracle helpers and the standalone CUDA checks.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
PKG       := multimodal-context-reasoning_b200
CSRC      := $(PKG)/csrc
LIB       := $(PKG)/libmodcr_b200.so
CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU_SRCS))
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)

all: $(LIB)

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Iinclude -c $< -o $@

$(LIB): $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $^

# developer build with the clock64() timeline of attn_bwd_kernel compiled in (scripts/attn_bwd_trace.py); never loaded by
# the package unless MODCR_B200_LIB points at it
TRACE_LIB  := $(PKG)/libmodcr_b200_trace.so
TRACE_OBJS := $(patsubst $(CSRC)/%.cu,build_trace/%.o,$(CU_SRCS))
build_trace/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build_trace
	$(NVCC) $(NVFLAGS) -DMODCR_ATTN_TRACE -Iinclude -c $< -o $@
$(TRACE_LIB): $(TRACE_OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $^
trace: $(TRACE_LIB)

cuda_tests: build/test_gemm

build/test_gemm: tests/cuda/test_gemm.cu build/gemm.o build/tma_host.o
	$(NVCC) $(NVFLAGS) -o $@ $^

clean:
	rm -rf build build_trace $(LIB) $(TRACE_LIB)

.PHONY: all clean cuda_tests trace
