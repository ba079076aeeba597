# Builds the product library (CUDA, sm_100a only) in-tree:  rchol_b200/lib/librchol_b200.so
# NB: -ccbin /usr/bin/g++ because the g++ first on PATH in this image links libstdc++ statically.
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  ?= /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -lineinfo -O3 -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall
SRC      := rchol_b200/csrc/rcg_api.cu rchol_b200/csrc/rcg_setup.cu rchol_b200/csrc/rcg_kernels.cu
OBJ      := $(SRC:.cu=.o)
LIB      := rchol_b200/lib/librchol_b200.so

all: $(LIB)

%.o: %.cu rchol_b200/csrc/rcg_common.cuh include/rchol_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p rchol_b200/lib
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $(OBJ) -cudart shared

clean:
	rm -f $(OBJ) $(LIB)
