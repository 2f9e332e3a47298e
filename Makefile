# Builds the C-ABI CUDA library (sm_100a only) in-tree so it travels to the GPU box.
PKG      := video-based-gait-analysis-for-dementia_b200
CSRC     := $(PKG)/csrc
OUT      := $(PKG)/lib/libgaitb200.so
NVCC     ?= nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Iinclude -I$(CSRC) --cudart static
SRCS     := $(wildcard $(CSRC)/*.cu)
OBJS     := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS     := $(wildcard $(CSRC)/*.cuh) include/gaitb200.h

all: $(OUT)

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(EXTRA) -c $< -o $@

$(OUT): $(OBJS)
	@mkdir -p $(PKG)/lib
	$(NVCC) $(ARCH) -shared --cudart static -o $@ $(OBJS)

clean:
	rm -rf build $(OUT)

.PHONY: all clean
