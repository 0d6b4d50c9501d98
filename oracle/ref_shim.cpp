// Builds the REFERENCE's own native code (header-only C++) into oracle/_ref/ -- TEST INFRASTRUCTURE.
// Nothing from the reference is copied into this repo: the header is included from where it lies
// (-I/root/reference/kiltro/FiniteElements, see oracle/Makefile).  The two extern "C" wrappers only
// forward to _ComputeSparsityPattern_ (ComputeSparsityPattern.h:22-62) and _ComputeDataIndices_
// (ComputeSparsityPattern.h:67-122) with the argument lists ComputeSparsityPattern.pyx:15-41 declares.
#include <cstdlib>
#include "ComputeSparsityPattern.h"

extern "C" int ref_compute_sparsity_pattern(const int *elements, const int *idx_start,
                                            const int *elem_container, int nvar, int nnode, int nelem,
                                            int nodeperelem, int idx_start_size, int *counts, int *indices) {
    int nnz = 0;
    _ComputeSparsityPattern_(elements, idx_start, elem_container, nvar, nnode, nelem, nodeperelem,
                             idx_start_size, counts, indices, nnz);
    return nnz;
}

extern "C" void ref_compute_data_indices(const int *indices, const int *indptr, int nelem, int nvar,
                                         int nodeperelem, const int *elements, const long *sorter,
                                         int *data_local_indices, int *data_global_indices) {
    _ComputeDataIndices_(indices, indptr, nelem, nvar, nodeperelem, elements, sorter,
                         data_local_indices, data_global_indices);
}
