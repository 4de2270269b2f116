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

cuda_tests: build/test_gemm

build/test_gemm: tests/cuda/test_gemm.cu build/gemm.o build/tma_host.o
	$(NVCC) $(NVFLAGS) -o $@ $^

clean:
	rm -rf build $(LIB)

.PHONY: all clean cuda_tests
