# Builds the product library (CUDA, sm_100a only) in-tree:  rchol_b200/lib/librchol_b200.so
# NB: -ccbin /usr/bin/g++ because the g++ first on PATH in this image links libstdc++ statically.
NVCC     ?= /usr/local/cuda/bin/nvcc
HOSTCXX  ?= /usr/bin/g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -lineinfo -O3 -std=c++17 -ccbin $(HOSTCXX) -Xcompiler -fPIC,-Wall,-fopenmp
SRC      := rchol_b200/csrc/rcg_api.cu rchol_b200/csrc/rcg_setup.cu rchol_b200/csrc/rcg_kernels.cu rchol_b200/csrc/rcg_trisolve.cu rchol_b200/csrc/rcg_blocked.cu rchol_b200/csrc/rcg_dist.cu rchol_b200/csrc/rcg_pool.cu
OBJ      := $(SRC:.cu=.o)
LIB      := rchol_b200/lib/librchol_b200.so

all: $(LIB) cxx

%.o: %.cu rchol_b200/csrc/rcg_common.cuh rchol_b200/csrc/rcg_device.cuh rchol_b200/csrc/rcg_cluster.cuh rchol_b200/csrc/rcg_fold.cuh rchol_b200/csrc/rcg_dense.cuh include/rchol_b200.h
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	mkdir -p rchol_b200/lib
	$(NVCC) $(ARCH) -ccbin $(HOSTCXX) -shared -o $@ $(OBJ) -cudart shared -ldl -lgomp

# C++ face of the drop-in (SparseCSR, pcg, util) and the example driver
CXXLIB   := rchol_b200/lib/librchol_b200_cxx.so
CXXSRC   := rchol_b200/cxx/sparse.cpp rchol_b200/cxx/pcg.cpp rchol_b200/cxx/util.cpp

$(CXXLIB): $(CXXSRC) $(LIB) rchol_b200/cxx/sparse.hpp rchol_b200/cxx/pcg.hpp rchol_b200/cxx/util.hpp
	$(HOSTCXX) -O2 -std=c++17 -fPIC -fopenmp -shared -o $@ $(CXXSRC) -Lrchol_b200/lib -lrchol_b200 -Wl,-rpath,'$$ORIGIN'

cxx: $(CXXLIB)

# needs baseline/_ref/librchol_producer.so (make -C baseline)
driver: rchol_b200/cxx/ex_laplace_parallel.cpp $(CXXLIB)
	$(HOSTCXX) -O2 -std=c++17 -o rchol_b200/lib/ex_laplace_parallel $< -Irchol_b200/cxx -Lrchol_b200/lib -lrchol_b200_cxx -lrchol_b200 \
	    -Lbaseline/_ref -lrchol_producer -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../../baseline/_ref'

# Drop-in proof: the reference's own example mains, UNMODIFIED and compiled where they lie (fed through stdin so that
# their `#include "sparse.hpp" / "util.hpp" / "pcg.hpp"` resolve to rchol_b200/cxx instead of the file's own directory),
# linked against our pcg class and the reference factorization.  Outputs into baseline/_ref (git-ignored, travels).
REF ?= /root/reference/c++
refmains: $(CXXLIB)
	for m in ex_laplace ex_laplace_parallel; do \
	  $(HOSTCXX) -O2 -std=c++17 -w -x c++ - -o baseline/_ref/ref_$$m -Irchol_b200/cxx -I$(REF)/rchol \
	    -Lrchol_b200/lib -lrchol_b200_cxx -lrchol_b200 -Lbaseline/_ref -lrchol_producer \
	    -Wl,-rpath,'$$ORIGIN' -Wl,-rpath,'$$ORIGIN/../../rchol_b200/lib' < $(REF)/$$m.cpp || exit 1; done

clean:
	rm -f $(OBJ) $(LIB) $(CXXLIB) rchol_b200/lib/ex_laplace_parallel baseline/_ref/ref_ex_laplace baseline/_ref/ref_ex_laplace_parallel
