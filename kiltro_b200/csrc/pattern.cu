// K2: symbolic phase -- CSR sparsity pattern and scatter/gather maps, entirely on device.
//
// Replaces kiltro/FiniteElements/ComputeSparsityPattern.pyx:44-112 (NumPy argsort/unique preparation + prefix sum)
// and ComputeSparsityPattern.h:22-62 (_ComputeSparsityPattern_: per-node sort+unique of neighbour lists) and
// :67-122 (_ComputeDataIndices_: per-entry binary search of the CSR slot).
//
// B200 formulation: every element entry (local row i, local col b; nodes visited in ascending node id as the
// reference does, .pyx:51-52) becomes a 64-bit key (row << cbits | col) with its entry id as payload; one stable LSD
// radix sort orders all entries by (row, col) and -- because the sort is stable and entry ids start ascending -- by
// ascending element inside each (row, col) group.  Group heads give the CSR pattern (rows strictly ascending, explicit
// zeros kept, exactly the reference's std::sort+std::unique result), the running head count gives
// data_global_indices, and the sorted payload is the deterministic row-owner gather map used by the numeric phase.
#include "common.cuh"
#include "scan.cuh"
#include "radix.cuh"

namespace {

// ----------------------------------------------------------------------------------------------- key generation
// one thread per element entry idx = e*cap + i*ndof + b  (ndof = npe*nvar, cap = ndof^2)
template <int NPE>
__global__ void __launch_bounds__(256) k_make_keys(const int32_t *__restrict__ elements, int64_t nent, int nvar,
                                                   int cbits, int64_t nnode, uint64_t *__restrict__ keys,
                                                   uint32_t *__restrict__ payload, int32_t *__restrict__ data_local,
                                                   int *__restrict__ bad) {
    const int ndof = NPE * nvar, cap = ndof * ndof;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool out_of_range = false;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < nent; idx += stride) {
        const int64_t e = idx / cap;
        const int r = (int)(idx - e * cap);
        const int i = r / ndof, b = r - i * ndof;
        // stable insertion argsort of the element's nodes ("sorter", .pyx:51)
        int node[NPE], sorter[NPE];
#pragma unroll
        for (int a = 0; a < NPE; ++a) {
            node[a] = __ldg(elements + e * NPE + a); sorter[a] = a;
            // a node id outside [0, nnode) would overflow the (row << cbits | col) key and send the head pass out of
            // bounds: clamp it (the result is discarded) and report KB_EINVAL
            if (node[a] < 0 || node[a] >= nnode) { out_of_range = true; node[a] = 0; }
        }
#pragma unroll
        for (int a = 1; a < NPE; ++a) {
#pragma unroll
            for (int c = a; c > 0; --c) {
                if (node[c - 1] > node[c]) {
                    int t = node[c]; node[c] = node[c - 1]; node[c - 1] = t;
                    t = sorter[c]; sorter[c] = sorter[c - 1]; sorter[c - 1] = t;
                }
            }
        }
        // DOF (sorted-node rank, component) of local row i and local column b                       .h:86-98
        const int in_ = i / nvar, ic = i - in_ * nvar, bn = b / nvar, bc = b - bn * nvar;
        int rn = 0, rs = 0, cn = 0, cs = 0;
#pragma unroll
        for (int a = 0; a < NPE; ++a) {        // register-friendly select instead of dynamic indexing
            if (a == in_) { rn = node[a]; rs = sorter[a]; }
            if (a == bn) { cn = node[a]; cs = sorter[a]; }
        }
        const uint64_t row = (uint64_t)rn * nvar + ic, col = (uint64_t)cn * nvar + bc;
        keys[idx] = (row << cbits) | col;
        payload[idx] = (uint32_t)idx;
        data_local[idx] = (rs * nvar + ic) * ndof + (cs * nvar + bc);                               // .h:114
    }
    if (out_of_range) *bad = 1;              // benign race: every writer stores the same value
}

// ----------------------------------------------------------------------------------------------- heads -> CSR
struct HeadIn {
    const uint64_t *keys;
    __device__ int operator()(int64_t p) const { return (p == 0 || keys[p] != keys[p - 1]) ? 1 : 0; }
};

struct HeadOut {
    const uint64_t *keys;
    const uint32_t *payload;
    const int32_t *data_local;
    int cbits, cap;
    int32_t *indptr, *indices, *data_global, *seg_ptr, *gather_src, *diag_pos;
    __device__ void operator()(int64_t p, int ex, int head) const {
        const int uid = ex + head - 1;                       // index of this entry's (row, col) group = CSR slot
        const uint32_t idx = payload[p];
        data_global[idx] = uid;                              // .h:113
        gather_src[p] = (int32_t)((idx / (uint32_t)cap) * (uint32_t)cap) + data_local[idx];
        if (head) {
            const uint64_t key = keys[p];
            const int64_t row = (int64_t)(key >> cbits);
            const int64_t col = (int64_t)(key & ((1ull << cbits) - 1ull));
            indices[uid] = (int32_t)col;
            seg_ptr[uid] = (int32_t)p;
            if (row == col) diag_pos[row] = uid;
            const int64_t prev_row = p == 0 ? -1 : (int64_t)(keys[p - 1] >> cbits);
            for (int64_t r = prev_row + 1; r <= row; ++r) indptr[r] = uid;      // also covers empty rows
        }
    }
};

__global__ void k_pattern_tail(const uint64_t *__restrict__ keys, int64_t nent, int cbits, int64_t nrows,
                               const int *__restrict__ total, int32_t *__restrict__ indptr,
                               int32_t *__restrict__ seg_ptr) {
    const int nnz = *total;
    const int64_t last_row = nent > 0 ? (int64_t)(keys[nent - 1] >> cbits) : -1;
    for (int64_t r = last_row + 1 + threadIdx.x; r <= nrows; r += blockDim.x) indptr[r] = nnz;
    if (threadIdx.x == 0) seg_ptr[nnz] = (int32_t)nent;
}

__global__ void __launch_bounds__(256) k_fill_i32(int32_t *__restrict__ p, int64_t n, int32_t v) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// CSR pattern of a quad-4 mesh numbered like Mesh.Rectangle (node (i,j) = j nxn + i), written down from the shape: row (i,j)
// holds the nodes of its <= 3 x 3 neighbourhood in ascending order -- exactly what the sort-based phase finds
// (test_pattern_grid_shortcut_bit_exact), without forming or sorting the 16 nelem element entries.
__global__ void __launch_bounds__(256) k_pattern_grid(int nxn, int nyn, int32_t *__restrict__ indptr,
                                                      int32_t *__restrict__ indices, int32_t *__restrict__ diag_pos) {
    const int64_t nn = (int64_t)nxn * nyn, stride = (int64_t)gridDim.x * blockDim.x;
    const int rowlen = 3 * nxn - 2;                       // entries of a mesh row of nodes, per neighbouring mesh row
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nn; p += stride) {
        const int i = (int)(p % nxn), j = (int)(p / nxn);
        const int cy = (j > 0) + 1 + (j < nyn - 1);
        const int before_y = j > 0 ? 3 * j - 1 : 0;       // sum of cy over the mesh rows below
        const int before_x = i > 0 ? 3 * i - 1 : 0;       // sum of cx over the nodes to the left
        int q = rowlen * before_y + cy * before_x;
        indptr[p] = q;
        for (int dj = (j > 0 ? -1 : 0); dj <= (j < nyn - 1 ? 1 : 0); ++dj)
            for (int di = (i > 0 ? -1 : 0); di <= (i < nxn - 1 ? 1 : 0); ++di) {
                if (dj == 0 && di == 0) diag_pos[p] = q;
                indices[q++] = (int32_t)(p + (int64_t)dj * nxn + di);
            }
        if (p == nn - 1) indptr[nn] = q;
    }
}

}  // namespace

extern "C" int kb_pattern_grid(int64_t nxn, int64_t nyn, int32_t *d_indptr, int32_t *d_indices, int32_t *d_diag_pos,
                               int64_t *h_nnz, void *stream) {
    if (!d_indptr || !d_indices || !d_diag_pos || !h_nnz || nxn < 2 || nyn < 2) return KB_EINVAL;
    const int64_t nnz = (3 * nxn - 2) * (3 * nyn - 2);
    if (nnz >= INT32_MAX || nxn * nyn >= INT32_MAX) return KB_ETOOBIG;
    k_pattern_grid<<<kb_grid(nxn * nyn, 256, 8), 256, 0, (cudaStream_t)stream>>>((int)nxn, (int)nyn, d_indptr, d_indices, d_diag_pos);
    KB_LAUNCH_CHECK();
    *h_nnz = nnz;
    return 0;
}

extern "C" int kb_pattern_build(const int32_t *d_elements, int64_t nelem, int npe, int64_t nnode, int nvar,
                                int32_t *d_indptr, int32_t *d_indices, int32_t *d_data_local, int32_t *d_data_global,
                                int32_t *d_seg_ptr, int32_t *d_gather_src, int32_t *d_diag_pos, int64_t *h_nnz,
                                void *stream) {
    if (!d_elements || !d_indptr || !d_indices || !d_data_local || !d_data_global || !d_seg_ptr || !d_gather_src ||
        !d_diag_pos || !h_nnz)
        return KB_EINVAL;
    if (nelem < 0 || nnode < 1 || nvar < 1 || (npe != 2 && npe != 4)) return KB_EINVAL;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t ndof = (int64_t)npe * nvar, cap = ndof * ndof, nent = nelem * cap, nrows = nnode * nvar;
    if (nent >= INT32_MAX || nrows >= INT32_MAX) return KB_ETOOBIG;
    const int cbits = kbradix::bits_for((uint64_t)(nrows - 1)), rbits = cbits;
    const int keybits = cbits + rbits;

    k_fill_i32<<<kb_grid(nrows, 256, 8), 256, 0, s>>>(d_diag_pos, nrows, -1);
    KB_LAUNCH_CHECK();
    if (nent == 0) {
        k_fill_i32<<<kb_grid(nrows + 1, 256, 8), 256, 0, s>>>(d_indptr, nrows + 1, 0);
        KB_LAUNCH_CHECK();
        k_fill_i32<<<1, 32, 0, s>>>(d_seg_ptr, 1, 0);
        KB_LAUNCH_CHECK();
        KB_CUDA(cudaStreamSynchronize(s));
        *h_nnz = 0;
        return 0;
    }

    // scratch: double-buffered keys/payload, per-tile histograms, scan spine -- released on EVERY exit path
    const int64_t nhist = kbradix::hist_ints(nent);
    const int64_t scan_ws = kbscan::workspace_ints(nent > nhist ? nent : nhist);
    struct Scratch {
        cudaStream_t s;
        void *p[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        explicit Scratch(cudaStream_t st) : s(st) {}
        ~Scratch() { for (void *q : p) if (q) cudaFreeAsync(q, s); }
    } scratch(s);
    const size_t bytes[7] = {sizeof(uint64_t) * (size_t)nent, sizeof(uint64_t) * (size_t)nent, sizeof(uint32_t) * (size_t)nent,
                             sizeof(uint32_t) * (size_t)nent, sizeof(int) * (size_t)nhist, sizeof(int) * (size_t)scan_ws,
                             sizeof(int) * 4};
    for (int k = 0; k < 7; ++k) KB_CUDA(cudaMallocAsync(&scratch.p[k], bytes[k], s));
    uint64_t *keys[2] = {(uint64_t *)scratch.p[0], (uint64_t *)scratch.p[1]};
    uint32_t *vals[2] = {(uint32_t *)scratch.p[2], (uint32_t *)scratch.p[3]};
    int *hist = (int *)scratch.p[4], *scanws = (int *)scratch.p[5], *bad = (int *)scratch.p[6];
    KB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int) * 4, s));

    const int g = kb_grid(nent, 256, 8);
    if (npe == 4) k_make_keys<4><<<g, 256, 0, s>>>(d_elements, nent, nvar, cbits, nnode, keys[0], vals[0], d_data_local, bad);
    else k_make_keys<2><<<g, 256, 0, s>>>(d_elements, nent, nvar, cbits, nnode, keys[0], vals[0], d_data_local, bad);
    KB_LAUNCH_CHECK();

    int cur = 0;
    {
        int rcs = kbradix::sort_pairs(keys, vals, nent, keybits, hist, scanws, s, &cur);
        if (rcs) return rcs;
    }

    HeadOut out{keys[cur], vals[cur], d_data_local, cbits, (int)cap,
                d_indptr, d_indices, d_data_global, d_seg_ptr, d_gather_src, d_diag_pos};
    int rc = kbscan::exclusive_scan(HeadIn{keys[cur]}, out, nent, scanws, s);
    if (rc) return rc;
    const int64_t nt = kbscan::num_tiles(nent);
    k_pattern_tail<<<1, 256, 0, s>>>(keys[cur], nent, cbits, nrows, scanws + nt, d_indptr, d_seg_ptr);
    KB_LAUNCH_CHECK();
    int nnz32 = 0, bad_h = 0;
    KB_CUDA(cudaMemcpyAsync(&nnz32, scanws + nt, sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaMemcpyAsync(&bad_h, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaStreamSynchronize(s));
    if (bad_h) return KB_EINVAL;             // connectivity refers to a node outside [0, nnode)
    *h_nnz = nnz32;
    return 0;
}
