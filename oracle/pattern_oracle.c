/* CPU oracle for the reference's one native component -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of kiltro/FiniteElements/ComputeSparsityPattern.pyx:44-112 (the NumPy
 * preparation) and ComputeSparsityPattern.h:22-62 (_ComputeSparsityPattern_) / :67-122
 * (_ComputeDataIndices_).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
 * may load the library built from this file; the product (kiltro_b200) never does.
 *
 * Parity status: PINNED against the reference's own outputs (tests/golden/pattern.npz) and, when
 * oracle/_ref/libkiltro_ref_pattern.so has been built from the reference header, against the
 * reference's native code directly (tests/test_oracle_golden.py).
 *
 * Build:  gcc -O2 -fPIC -shared pattern_oracle.c -o _build/libkiltro_oracle.so   (oracle/Makefile)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/* position of val in the ascending array a[0..n) (val is known to be present) */
static int32_t locate(const int32_t *a, int32_t n, int32_t val) {
    int32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        int32_t mid = lo + (hi - lo) / 2;
        if (val < a[mid]) hi = mid; else lo = mid;
    }
    return lo;
}

/* Pass 1 (indptr) + pass 2 (indices): rows = DOF, row of node n lists nvar*m+l for every node m that
 * shares an element with n, ascending, explicit zeros kept (.h:36-61; prefix sum .pyx:86-89).
 * elements: (nelem, npe) int64 as in Mesh.elements.  indptr: (nnode*nvar+1).  indices: capacity
 * (npe*nvar)^2*nelem (.pyx:64).  Returns nnz, or -1 on allocation failure. */
int64_t oracle_sparsity_pattern(const int64_t *elements, int64_t nelem, int32_t npe, int64_t nnode,
                                int32_t nvar, int32_t *indptr, int32_t *indices) {
    int64_t ninc = nelem * npe, i, e;
    int32_t *deg = (int32_t *)calloc((size_t)nnode + 1, sizeof(int32_t));
    int32_t *adj = (int32_t *)malloc((size_t)(ninc > 0 ? ninc : 1) * sizeof(int32_t));
    int32_t *cur = (int32_t *)malloc((size_t)(nnode + 1) * sizeof(int32_t));
    if (!deg || !adj || !cur) return -1;
    /* node -> incident elements, ascending element id (what argsort/unique build in .pyx:55-60) */
    for (i = 0; i < ninc; ++i) deg[elements[i] + 1]++;
    for (i = 0; i < nnode; ++i) deg[i + 1] += deg[i];
    memcpy(cur, deg, (size_t)(nnode + 1) * sizeof(int32_t));
    for (e = 0; e < nelem; ++e)
        for (i = 0; i < npe; ++i) adj[cur[elements[e * npe + i]]++] = (int32_t)e;
    int64_t nnz = 0;
    int32_t cap = 64, *loc = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    indptr[0] = 0;
    for (i = 0; i < nnode; ++i) {
        int32_t m = (deg[i + 1] - deg[i]) * npe, c = 0, u = 0, j, k, l;
        if (m > cap) { cap = 2 * m; loc = (int32_t *)realloc(loc, sizeof(int32_t) * (size_t)cap); }
        for (j = deg[i]; j < deg[i + 1]; ++j)
            for (k = 0; k < npe; ++k) loc[c++] = (int32_t)elements[(int64_t)adj[j] * npe + k];
        qsort(loc, (size_t)c, sizeof(int32_t), cmp_i32);          /* std::sort   .h:46 */
        for (j = 0; j < c; ++j) if (j == 0 || loc[j] != loc[j - 1]) loc[u++] = loc[j];   /* std::unique */
        for (j = 0; j < nvar; ++j) {                              /* .h:51-59 */
            for (k = 0; k < u; ++k)
                for (l = 0; l < nvar; ++l) indices[nnz++] = nvar * loc[k] + l;
            int64_t len = (int64_t)u * nvar;                      /* .pyx:88 */
            if (len > nnode * nvar) len = nnode * nvar;
            indptr[i * nvar + j + 1] = indptr[i * nvar + j] + (int32_t)len;
        }
    }
    free(loc); free(cur); free(adj); free(deg);
    return nnz;
}

/* data_global_indices / data_local_indices (.h:67-122): element nodes are visited in ascending node id
 * ("sorter" = argsort of the element's nodes, .pyx:51-52); local index refers to the unsorted, raveled
 * element matrix. */
void oracle_data_indices(const int64_t *elements, int64_t nelem, int32_t npe, int32_t nvar,
                         const int32_t *indptr, const int32_t *indices,
                         int32_t *data_local, int32_t *data_global) {
    int32_t ndof = npe * nvar, i, j, a, b;
    int64_t e, cap = (int64_t)ndof * ndof;
    int32_t *srt = (int32_t *)malloc(sizeof(int32_t) * (size_t)npe);
    int32_t *sorter = (int32_t *)malloc(sizeof(int32_t) * (size_t)npe);
    int32_t *gcol = (int32_t *)malloc(sizeof(int32_t) * (size_t)ndof);
    int32_t *lcol = (int32_t *)malloc(sizeof(int32_t) * (size_t)ndof);
    for (e = 0; e < nelem; ++e) {
        for (i = 0; i < npe; ++i) { sorter[i] = i; }
        for (i = 1; i < npe; ++i) {                               /* stable insertion argsort */
            int32_t s = sorter[i]; int64_t v = elements[e * npe + s];
            for (j = i - 1; j >= 0 && elements[e * npe + sorter[j]] > v; --j) sorter[j + 1] = sorter[j];
            sorter[j + 1] = s;
        }
        for (i = 0; i < npe; ++i) srt[i] = (int32_t)elements[e * npe + sorter[i]];
        for (i = 0; i < npe; ++i)
            for (a = 0; a < nvar; ++a) {
                gcol[nvar * i + a] = nvar * srt[i] + a;           /* .h:86-91 */
                lcol[nvar * i + a] = nvar * sorter[i] + a;        /* .h:93-98 */
            }
        for (i = 0; i < ndof; ++i) {
            int32_t row = gcol[i], r0 = indptr[row], len = indptr[row + 1] - r0;
            for (b = 0; b < ndof; ++b) {
                data_global[e * cap + (int64_t)i * ndof + b] = r0 + locate(indices + r0, len, gcol[b]);
                data_local[e * cap + (int64_t)i * ndof + b] = lcol[i] * ndof + lcol[b];
            }
        }
    }
    free(srt); free(sorter); free(gcol); free(lcol);
}
