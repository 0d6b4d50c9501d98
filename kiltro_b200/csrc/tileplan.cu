// Symbolic phase, part 2: the tile plan of the fused assembly kernel (layout in tileplan.cuh), built on device from the
// CSR pattern and the gather map kb_pattern_build produced.  One-off per mesh, cached with the pattern
// (FEMSolver.ComputeSparsityFEM, mirroring the reference's is_sparsity_pattern_computed: FEMSolver.py:315-324).
#include "common.cuh"
#include "radix.cuh"
#include "scan.cuh"
#include "tileplan.cuh"

using namespace kbtile;

namespace {

__global__ void __launch_bounds__(256) k_tile_rows(Geom g, int32_t *__restrict__ tile_rows,
                                                   int32_t *__restrict__ tile_row_ptr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.nnode; i += stride) {
        int t, slot;
        tile_of(g, (int)i, t, slot);
        tile_rows[slot] = (int)i;
    }
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t <= g.ntiles; t += stride)
        tile_row_ptr[t] = tile_row_start(g, (int)t);
}

// one item per (element, local node): key = tile of the node if this is the first local node of the element lying in
// that tile, else the sentinel ntiles; payload = e*NPE + a.  Items start in ascending element order and the sort is
// stable, so each tile's element list comes out ascending.
template <int NPE>
__global__ void __launch_bounds__(256) k_tile_items(Geom g, const int32_t *__restrict__ elements, int64_t nelem,
                                                    uint64_t *__restrict__ keys, uint32_t *__restrict__ payload) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nelem; e += stride) {
        int tl[NPE];
#pragma unroll
        for (int a = 0; a < NPE; ++a) {
            int slot;
            tile_of(g, __ldg(elements + e * NPE + a), tl[a], slot);
        }
#pragma unroll
        for (int a = 0; a < NPE; ++a) {
            bool first = true;
#pragma unroll
            for (int b = 0; b < NPE; ++b)
                if (b < a && tl[b] == tl[a]) first = false;
            keys[e * NPE + a] = first ? (uint64_t)tl[a] : (uint64_t)g.ntiles;
            payload[e * NPE + a] = (uint32_t)(e * NPE + a);
        }
    }
}

__global__ void __launch_bounds__(256) k_tile_elems(int ntiles, int npe, const uint64_t *__restrict__ keys,
                                                    const uint32_t *__restrict__ payload, int64_t nitems,
                                                    const int32_t *__restrict__ elements,
                                                    int32_t *__restrict__ tile_elems, int32_t *__restrict__ tile_conn,
                                                    int32_t *__restrict__ tile_elem_ptr) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nitems; p += stride) {
        const int key = (int)keys[p];
        if (key < ntiles) {
            const int e = (int)(payload[p] / (uint32_t)npe);
            tile_elems[p] = e;
            for (int b = 0; b < npe; ++b) tile_conn[p * npe + b] = __ldg(elements + (int64_t)e * npe + b);   // tile-local copy
        }
        const int prev = p == 0 ? -1 : (int)keys[p - 1];
        for (int t = prev + 1; t <= key && t <= ntiles; ++t) tile_elem_ptr[t] = (int32_t)p;
        if (p == nitems - 1)
            for (int t = key + 1; t <= ntiles; ++t) tile_elem_ptr[t] = (int32_t)nitems;     // no sentinel items at all
    }
}

// number of incident elements of the row at tile-order position j (length of the diagonal entry's gather segment)
struct DegIn {
    const int32_t *tile_rows, *diag_pos, *seg_ptr;
    __device__ int operator()(int64_t j) const {
        const int d = diag_pos[tile_rows[j]];
        return d >= 0 ? seg_ptr[d + 1] - seg_ptr[d] : 0;
    }
};
struct DegOut {
    int32_t *pair_ptr;
    __device__ void operator()(int64_t j, int ex, int) const { pair_ptr[j] = ex; }
};
__global__ void k_set_total(const int *total, int32_t *dst) { *dst = *total; }

__device__ __forceinline__ int lower_bound_i32(const int32_t *a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int NPE>
__global__ void __launch_bounds__(256) k_pair_rec(Geom g, const int32_t *__restrict__ elements,
                                                  const int32_t *__restrict__ indptr,
                                                  const int32_t *__restrict__ indices,
                                                  const int32_t *__restrict__ seg_ptr,
                                                  const int32_t *__restrict__ gather_src,
                                                  const int32_t *__restrict__ diag_pos,
                                                  const int32_t *__restrict__ tile_rows,
                                                  const int32_t *__restrict__ tile_elems,
                                                  const int32_t *__restrict__ tile_elem_ptr,
                                                  const int32_t *__restrict__ pair_ptr,
                                                  uint32_t *__restrict__ pair_rec) {
    constexpr int CAP = NPE * NPE;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < g.nnode; j += stride) {
        const int row = tile_rows[j];
        int t, slot_unused;
        tile_of(g, row, t, slot_unused);
        const int d = diag_pos[row];
        if (d < 0) continue;
        const int eb = tile_elem_ptr[t], ne = tile_elem_ptr[t + 1] - eb;
        const int p0 = indptr[row], len = indptr[row + 1] - p0;
        int out = pair_ptr[j];
        for (int p = seg_ptr[d], pe = seg_ptr[d + 1]; p < pe; ++p, ++out) {
            const int src = gather_src[p];
            const int e = src / CAP, a = (src - e * CAP) / (NPE + 1);
            const int slot = lower_bound_i32(tile_elems + eb, ne, e);
            uint32_t rec = ((uint32_t)slot << 18) | ((uint32_t)a << 16);
            bool wide = false;
#pragma unroll
            for (int b = 0; b < NPE; ++b) {
                const int k = lower_bound_i32(indices + p0, len, __ldg(elements + (int64_t)e * NPE + b));
                if (k >= 16) wide = true;
                rec |= (uint32_t)(k & 15) << (4 * b);
            }
            if (wide) rec |= 0x80000000u;
            pair_rec[out] = rec;
        }
    }
}

// stats[0] = max elements per tile, [1] = max rows per tile, [2] = max CSR entries per tile
__global__ void __launch_bounds__(256) k_tile_stats(int ntiles, const int32_t *__restrict__ tile_row_ptr,
                                                    const int32_t *__restrict__ tile_rows,
                                                    const int32_t *__restrict__ tile_elem_ptr,
                                                    const int32_t *__restrict__ indptr, int *__restrict__ stats) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += stride) {
        const int rb = tile_row_ptr[t], re = tile_row_ptr[t + 1];
        int nv = 0;
        for (int j = rb; j < re; ++j) { const int r = tile_rows[j]; nv += indptr[r + 1] - indptr[r]; }
        atomicMax(stats + 0, tile_elem_ptr[t + 1] - tile_elem_ptr[t]);
        atomicMax(stats + 1, re - rb);
        atomicMax(stats + 2, nv);
    }
}

}  // namespace

extern "C" int kb_tile_plan_sizes(int64_t nnode, int64_t nelem, int npe, int64_t nx_nodes, int64_t ny_nodes,
                                  int64_t *h_sizes) {
    if (!h_sizes || nnode < 1 || nelem < 0 || (npe != 2 && npe != 4)) return KB_EINVAL;
    const Geom g = make_geom(nnode, npe, nx_nodes, ny_nodes);
    h_sizes[0] = g.ntiles;
    h_sizes[1] = nelem * npe;      // capacity of tile_elems (upper bound; the exact count is returned by the build)
    h_sizes[2] = nelem * npe;      // capacity of pair_rec
    h_sizes[3] = g.structured;
    return 0;
}

extern "C" int kb_tile_plan_build(const int32_t *d_elements, int64_t nelem, int npe, int64_t nnode,
                                  const int32_t *d_indptr, const int32_t *d_indices, const int32_t *d_seg_ptr,
                                  const int32_t *d_gather_src, const int32_t *d_diag_pos, int64_t nx_nodes,
                                  int64_t ny_nodes, int32_t *d_tile_rows, int32_t *d_tile_row_ptr,
                                  int32_t *d_tile_elems, int32_t *d_tile_conn, int32_t *d_tile_elem_ptr,
                                  uint32_t *d_pair_rec, int32_t *d_pair_ptr, int64_t *h_info, void *stream) {
    if (!d_elements || !d_indptr || !d_indices || !d_seg_ptr || !d_gather_src || !d_diag_pos || !d_tile_rows ||
        !d_tile_row_ptr || !d_tile_elems || !d_tile_conn || !d_tile_elem_ptr || !d_pair_rec || !d_pair_ptr || !h_info)
        return KB_EINVAL;
    if (nnode < 1 || nelem < 1 || (npe != 2 && npe != 4)) return KB_EINVAL;
    if (nelem * npe >= INT32_MAX) return KB_ETOOBIG;
    cudaStream_t s = (cudaStream_t)stream;
    const Geom g = make_geom(nnode, npe, nx_nodes, ny_nodes);
    const int64_t nitems = nelem * npe;

    k_tile_rows<<<kb_grid(nnode, 256, 8), 256, 0, s>>>(g, d_tile_rows, d_tile_row_ptr);
    KB_LAUNCH_CHECK();

    const int64_t nhist = kbradix::hist_ints(nitems);
    const int64_t big = nitems > nnode ? nitems : nnode;
    const int64_t scan_ws = kbscan::workspace_ints(big > nhist ? big : nhist);
    uint64_t *keys[2];
    uint32_t *vals[2];
    int *hist, *scanws, *stats;
    KB_CUDA(cudaMallocAsync((void **)&keys[0], sizeof(uint64_t) * nitems, s));
    KB_CUDA(cudaMallocAsync((void **)&keys[1], sizeof(uint64_t) * nitems, s));
    KB_CUDA(cudaMallocAsync((void **)&vals[0], sizeof(uint32_t) * nitems, s));
    KB_CUDA(cudaMallocAsync((void **)&vals[1], sizeof(uint32_t) * nitems, s));
    KB_CUDA(cudaMallocAsync((void **)&hist, sizeof(int) * nhist, s));
    KB_CUDA(cudaMallocAsync((void **)&scanws, sizeof(int) * scan_ws, s));
    KB_CUDA(cudaMallocAsync((void **)&stats, sizeof(int) * 4, s));
    KB_CUDA(cudaMemsetAsync(stats, 0, sizeof(int) * 4, s));

    const int ge = kb_grid(nelem, 256, 8);
    if (npe == 4) k_tile_items<4><<<ge, 256, 0, s>>>(g, d_elements, nelem, keys[0], vals[0]);
    else k_tile_items<2><<<ge, 256, 0, s>>>(g, d_elements, nelem, keys[0], vals[0]);
    KB_LAUNCH_CHECK();
    int cur = 0;
    int rc = kbradix::sort_pairs(keys, vals, nitems, kbradix::bits_for((uint64_t)g.ntiles), hist, scanws, s, &cur);
    if (rc) return rc;
    k_tile_elems<<<kb_grid(nitems, 256, 8), 256, 0, s>>>(g.ntiles, npe, keys[cur], vals[cur], nitems, d_elements,
                                                        d_tile_elems, d_tile_conn, d_tile_elem_ptr);
    KB_LAUNCH_CHECK();

    rc = kbscan::exclusive_scan(DegIn{d_tile_rows, d_diag_pos, d_seg_ptr}, DegOut{d_pair_ptr}, nnode, scanws, s);
    if (rc) return rc;
    k_set_total<<<1, 1, 0, s>>>(scanws + kbscan::num_tiles(nnode), d_pair_ptr + nnode);
    KB_LAUNCH_CHECK();

    const int gn = kb_grid(nnode, 256, 8);
    if (npe == 4)
        k_pair_rec<4><<<gn, 256, 0, s>>>(g, d_elements, d_indptr, d_indices, d_seg_ptr, d_gather_src, d_diag_pos,
                                         d_tile_rows, d_tile_elems, d_tile_elem_ptr, d_pair_ptr, d_pair_rec);
    else
        k_pair_rec<2><<<gn, 256, 0, s>>>(g, d_elements, d_indptr, d_indices, d_seg_ptr, d_gather_src, d_diag_pos,
                                         d_tile_rows, d_tile_elems, d_tile_elem_ptr, d_pair_ptr, d_pair_rec);
    KB_LAUNCH_CHECK();
    k_tile_stats<<<kb_grid(g.ntiles, 256, 8), 256, 0, s>>>(g.ntiles, d_tile_row_ptr, d_tile_rows, d_tile_elem_ptr,
                                                          d_indptr, stats);
    KB_LAUNCH_CHECK();

    int h_stats[4] = {0, 0, 0, 0}, h_nte = 0, h_np = 0;
    KB_CUDA(cudaMemcpyAsync(h_stats, stats, sizeof(int) * 4, cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaMemcpyAsync(&h_nte, d_tile_elem_ptr + g.ntiles, sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaMemcpyAsync(&h_np, d_pair_ptr + nnode, sizeof(int), cudaMemcpyDeviceToHost, s));
    KB_CUDA(cudaFreeAsync(keys[0], s)); KB_CUDA(cudaFreeAsync(keys[1], s));
    KB_CUDA(cudaFreeAsync(vals[0], s)); KB_CUDA(cudaFreeAsync(vals[1], s));
    KB_CUDA(cudaFreeAsync(hist, s)); KB_CUDA(cudaFreeAsync(scanws, s)); KB_CUDA(cudaFreeAsync(stats, s));
    KB_CUDA(cudaStreamSynchronize(s));
    h_info[0] = g.ntiles; h_info[1] = h_nte; h_info[2] = h_np;
    h_info[3] = h_stats[0]; h_info[4] = h_stats[1]; h_info[5] = h_stats[2];
    h_info[6] = (h_stats[0] <= EMAX && h_stats[1] <= RMAX && h_stats[2] <= VMAX) ? 1 : 0;   // fits the fused kernel
    h_info[7] = g.structured;
    return 0;
}
