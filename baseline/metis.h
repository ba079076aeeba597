// Minimal stand-in for METIS' public header: the CUDA 12.9 toolkit ships
// libmetis_static.a (64-bit idx_t) but no metis.h.  Only the single entry point
// the reference calls (/root/reference/c++/rchol/find_separator.cpp:90) is declared.
#pragma once
#include <stdint.h>
typedef int64_t idx_t;
#ifdef __cplusplus
extern "C" {
#endif
int METIS_ComputeVertexSeparator(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt,
                                 idx_t *options, idx_t *sepsize, idx_t *part);
#ifdef __cplusplus
}
#endif
