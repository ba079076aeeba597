/* Stand-in for Intel MKL's mkl_types.h -- ORACLE BUILD ONLY (see oracle/Makefile).
 * The reference's pcg.cpp includes mkl_spblas.h / mkl.h / mkl_types.h with `#define MKL_INT size_t`
 * (ILP64; /root/reference/c++/util/pcg.hpp:6-9).  No MKL development package exists in this image, but
 * libtorch_cpu.so exports the genuine oneMKL 2024.2 LP64 kernels.  These headers declare exactly the
 * names pcg.cpp uses and route them to adapters (oracle/mkl_adapter.cpp) that narrow the 64-bit index
 * arrays and call the real LP64 entry points. */
#ifndef RCHOL_B200_MKLSHIM_TYPES_H
#define RCHOL_B200_MKLSHIM_TYPES_H
#include <stddef.h>
#ifndef MKL_INT
#define MKL_INT size_t
#endif
#endif
