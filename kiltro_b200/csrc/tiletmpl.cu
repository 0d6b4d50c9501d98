// Symbolic phase, part 3: tile templates of the fused assembly (layout and rationale in tileplan.cuh).  Built on device
// from the tile plan; one-off per mesh, cached with the pattern like the rest of the symbolic data
// (FEMSolver.ComputeSparsityFEM, mirroring FEMSolver.py:313-325).
//   pass 1  one CTA per tile: derive the tile's local plan in shared memory, hash it, count its runs
//   host    distinct hashes -> class ids (<= CMAX), one representative tile per class, prefix sum of the run counts
//   pass 2  one CTA per class: derive the representative's plan again and store it as the template
//   pass 3  one CTA per tile: derive again, compare word by word with the template of its class (a hash collision or any
//           irregularity clears the `ok` flag -> the caller uses the generic kernel) and write the first row of each run
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tileplan.cuh"

using namespace kbtile;

namespace {

constexpr int TWORDS = (int)(sizeof(TileTmpl) / 4);

struct PlanView {
    const int32_t *indptr, *tile_rows, *tile_row_ptr, *tile_elem_ptr, *pair_ptr;
    const uint32_t *pair_rec;
};

// CTA-cooperative (256 threads).  Returns false (uniformly) when the tile cannot be expressed as a template.
__device__ bool derive(const PlanView &pv, int t, TileTmpl *img, uint8_t *cnt, int *sScan, int *sFail) {
    const int tid = threadIdx.x;
    const int rb = pv.tile_row_ptr[t], nr = pv.tile_row_ptr[t + 1] - rb;
    const int ne = pv.tile_elem_ptr[t + 1] - pv.tile_elem_ptr[t];
    uint32_t *w = reinterpret_cast<uint32_t *>(img);
    for (int i = tid; i < TWORDS; i += THREADS) w[i] = 0u;
    if (tid == 0) *sFail = 0;
    __syncthreads();
    if (nr > RMAX || ne > EMAX) return false;
    for (int i = tid; i < VMAX * 4; i += THREADS) (&img->ent_src[0][0])[i] = (uint16_t)(KZERO * 8);
    for (int i = tid; i < RMAX * 4; i += THREADS) (&img->row_src[0][0])[i] = (uint16_t)(FZERO * 8);
    for (int i = tid; i < VMAX; i += THREADS) cnt[i] = 0;
    int row = 0, rowp = -2, p0 = 0, len = 0;
    if (tid < nr) {
        row = pv.tile_rows[rb + tid];
        if (tid > 0) rowp = pv.tile_rows[rb + tid - 1];
        p0 = pv.indptr[row];
        len = pv.indptr[row + 1] - p0;
    }
    const int segstart = (tid < nr && row != rowp + 1) ? 1 : 0;
    int packed_total;
    const bool big = len > 0xfff;                               // keeps the packed scan exact
    const int ex = block_scan_excl((big ? 0xfff : len) | (segstart << 16), sScan, &packed_total);
    const int off = ex & 0xffff, total = packed_total & 0xffff, nseg = packed_total >> 16;
    const int run = (ex >> 16) + segstart - 1;
    if (big) *sFail = 1;
    __syncthreads();
    if (*sFail || total > VMAX || nseg > RUNMAX) return false;
    if (segstart) { img->run_ent[run] = (uint16_t)off; img->run_row[run] = (uint16_t)tid; }
    if (tid == 0) {
        img->run_ent[nseg] = (uint16_t)total; img->run_row[nseg] = (uint16_t)nr;
        img->nent = total; img->nrow = nr; img->nrun = nseg; img->ne = ne;
    }
    if (tid < nr) {
        img->row_run[tid] = (uint8_t)run;
        for (int i = 0; i < len; ++i) img->ent_run[off + i] = (uint8_t)run;
        const int jp0 = pv.pair_ptr[rb + tid], jp1 = pv.pair_ptr[rb + tid + 1];
        if (jp1 - jp0 > 4) *sFail = 1;
        for (int j = jp0; j < jp1 && j < jp0 + 4; ++j) {
            const uint32_t rec = pv.pair_rec[j];
            const int slot = (int)((rec >> 18) & 0x1fffu), a = (int)((rec >> 16) & 3u);
            if ((rec & 0x80000000u) || slot >= EMAX) { *sFail = 1; break; }
            img->row_src[tid][j - jp0] = (uint16_t)((a * FSTRIDE + slot) * 8);
            for (int b = 0; b < 4; ++b) {
                const int pos = (int)((rec >> (4 * b)) & 15u);
                if (pos >= len) { *sFail = 1; break; }
                const int e = off + pos, c = cnt[e];
                if (c >= 4) { *sFail = 1; break; }
                img->ent_src[e][c] = (uint16_t)((((a * 4 + b) * KSTRIDE) + slot) * 8);
                cnt[e] = (uint8_t)(c + 1);
            }
        }
    }
    __syncthreads();
    return *sFail == 0;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

struct Smem {
    TileTmpl img;
    uint8_t cnt[VMAX];
    int scan[64];
    int fail;
    unsigned long long hash;
};

// mode 0: hash + run count;  mode 1: store as template (tile list = representatives);  mode 2: verify + run rows
__global__ void __launch_bounds__(THREADS) k_tmpl_pass(PlanView pv, int mode, int ntiles, const int32_t *__restrict__ reps,
                                                       unsigned long long *__restrict__ hashes, int32_t *__restrict__ nruns,
                                                       TileTmpl *__restrict__ templates,
                                                       const uint8_t *__restrict__ tile_class,
                                                       const int32_t *__restrict__ trun_ptr, int32_t *__restrict__ trun_row,
                                                       int *__restrict__ bad) {
    extern __shared__ __align__(16) unsigned char raw[];
    Smem *sm = reinterpret_cast<Smem *>(raw);
    const int tid = threadIdx.x;
    for (int item = blockIdx.x; item < ntiles; item += gridDim.x) {
        const int t = mode == 1 ? reps[item] : item;
        const bool ok = derive(pv, t, &sm->img, sm->cnt, sm->scan, &sm->fail);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(&sm->img);
        if (mode == 0) {
            if (tid == 0) sm->hash = 0ull;
            __syncthreads();
            unsigned long long h = 0ull;
            for (int i = tid; i < TWORDS; i += THREADS) h += mix64(((uint64_t)w[i] << 32) | (uint32_t)(i * 2654435761u + 1u));
            for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(KB_FULL, h, o);
            if ((tid & 31) == 0) atomicAdd(&sm->hash, h);
            __syncthreads();
            if (tid == 0) {
                hashes[t] = ok ? (sm->hash | 1ull) : 0ull;          // 0 = not expressible
                nruns[t] = ok ? sm->img.nrun : 0;
            }
        } else if (mode == 1) {
            uint32_t *dst = reinterpret_cast<uint32_t *>(templates + item);
            for (int i = tid; i < TWORDS; i += THREADS) dst[i] = w[i];
            if (!ok && tid == 0) atomicExch(bad, 1);
        } else {
            const uint32_t *ref = reinterpret_cast<const uint32_t *>(templates + tile_class[t]);
            int diff = ok ? 0 : 1;
            for (int i = tid; i < TWORDS; i += THREADS) diff |= (w[i] != ref[i]);
            if (diff) atomicExch(bad, 1);
            const int rb = pv.tile_row_ptr[t], base = trun_ptr[t];
            if (ok && tid < sm->img.nrun) trun_row[base + tid] = pv.tile_rows[rb + sm->img.run_row[tid]];
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" int kb_tile_tmpl_sizes(int64_t *h_sizes) {
    if (!h_sizes) return KB_EINVAL;
    h_sizes[0] = (int64_t)sizeof(TileTmpl);
    h_sizes[1] = CMAX;
    h_sizes[2] = RUNMAX;
    return 0;
}

extern "C" int kb_tile_tmpl_build(int npe, int64_t nnode, int64_t ntiles, const int32_t *d_indptr,
                                  const int32_t *d_tile_rows, const int32_t *d_tile_row_ptr,
                                  const int32_t *d_tile_elem_ptr, const uint32_t *d_pair_rec,
                                  const int32_t *d_pair_ptr, uint8_t *d_tile_class, int32_t *d_trun_ptr,
                                  int32_t *d_trun_row, void *d_templates, int64_t *h_info, void *stream) {
    if (!d_indptr || !d_tile_rows || !d_tile_row_ptr || !d_tile_elem_ptr || !d_pair_rec || !d_pair_ptr ||
        !d_tile_class || !d_trun_ptr || !d_trun_row || !d_templates || !h_info)
        return KB_EINVAL;
    h_info[0] = 0; h_info[1] = 0; h_info[2] = 0;
    if (npe != 4 || ntiles < 1 || nnode < 1) return 0;               // templates serve the quad kernel only
    cudaStream_t s = (cudaStream_t)stream;
    static KbOncePerDevice configured;
    if (configured.need()) {
        KB_CUDA(cudaFuncSetAttribute(k_tmpl_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
    }
    const PlanView pv{d_indptr, d_tile_rows, d_tile_row_ptr, d_tile_elem_ptr, d_pair_ptr, d_pair_rec};
    unsigned long long *d_hash;
    int32_t *d_nruns, *d_reps;
    int *d_bad;
    KB_CUDA(cudaMallocAsync((void **)&d_hash, sizeof(unsigned long long) * ntiles, s));
    KB_CUDA(cudaMallocAsync((void **)&d_nruns, sizeof(int32_t) * ntiles, s));
    KB_CUDA(cudaMallocAsync((void **)&d_reps, sizeof(int32_t) * CMAX, s));
    KB_CUDA(cudaMallocAsync((void **)&d_bad, sizeof(int), s));
    KB_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), s));
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)kb_num_sms() * 2);
    int rc = 0;
    std::vector<unsigned long long> h_hash((size_t)ntiles);
    std::vector<int32_t> h_nruns((size_t)ntiles), h_ptr((size_t)ntiles + 1), h_reps;
    std::vector<uint8_t> h_class((size_t)ntiles);
    bool ok = true;
    do {
        k_tmpl_pass<<<grid, THREADS, sizeof(Smem), s>>>(pv, 0, (int)ntiles, nullptr, d_hash, d_nruns, nullptr, nullptr,
                                                       nullptr, nullptr, d_bad);
        KB_COUNT_LAUNCH();
        if ((rc = (int)cudaGetLastError())) break;
        if ((rc = (int)cudaMemcpyAsync(h_hash.data(), d_hash, sizeof(unsigned long long) * ntiles, cudaMemcpyDeviceToHost, s))) break;
        if ((rc = (int)cudaMemcpyAsync(h_nruns.data(), d_nruns, sizeof(int32_t) * ntiles, cudaMemcpyDeviceToHost, s))) break;
        if ((rc = (int)cudaStreamSynchronize(s))) break;
        // distinct hashes -> classes, in order of first appearance
        std::vector<std::pair<unsigned long long, int>> order((size_t)ntiles);
        for (int64_t t = 0; t < ntiles; ++t) order[(size_t)t] = {h_hash[(size_t)t], (int)t};
        std::sort(order.begin(), order.end());
        int64_t total_runs = 0;
        for (size_t i = 0; i < order.size() && ok; ++i) {
            if (order[i].first == 0ull) { ok = false; break; }
            if (i == 0 || order[i].first != order[i - 1].first) {
                if ((int)h_reps.size() == CMAX) { ok = false; break; }
                h_reps.push_back(order[i].second);                   // smallest tile id with this hash
            }
            h_class[(size_t)order[i].second] = (uint8_t)(h_reps.size() - 1);
        }
        if (!ok) break;
        for (int64_t t = 0; t < ntiles; ++t) { h_ptr[(size_t)t] = (int32_t)total_runs; total_runs += h_nruns[(size_t)t]; }
        h_ptr[(size_t)ntiles] = (int32_t)total_runs;
        if (total_runs > nnode) { ok = false; break; }
        if ((rc = (int)cudaMemcpyAsync(d_tile_class, h_class.data(), (size_t)ntiles, cudaMemcpyHostToDevice, s))) break;
        if ((rc = (int)cudaMemcpyAsync(d_trun_ptr, h_ptr.data(), sizeof(int32_t) * (ntiles + 1), cudaMemcpyHostToDevice, s))) break;
        if ((rc = (int)cudaMemcpyAsync(d_reps, h_reps.data(), sizeof(int32_t) * h_reps.size(), cudaMemcpyHostToDevice, s))) break;
        const int nc = (int)h_reps.size();
        k_tmpl_pass<<<nc, THREADS, sizeof(Smem), s>>>(pv, 1, nc, d_reps, nullptr, nullptr, (TileTmpl *)d_templates, nullptr,
                                                     nullptr, nullptr, d_bad);
        KB_COUNT_LAUNCH();
        if ((rc = (int)cudaGetLastError())) break;
        k_tmpl_pass<<<grid, THREADS, sizeof(Smem), s>>>(pv, 2, (int)ntiles, nullptr, nullptr, nullptr,
                                                       (TileTmpl *)d_templates, d_tile_class, d_trun_ptr, d_trun_row, d_bad);
        KB_COUNT_LAUNCH();
        if ((rc = (int)cudaGetLastError())) break;
        int h_bad = 0;
        if ((rc = (int)cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, s))) break;
        if ((rc = (int)cudaStreamSynchronize(s))) break;
        if (h_bad) ok = false;
        h_info[1] = nc; h_info[2] = total_runs;
    } while (0);
    cudaFreeAsync(d_hash, s); cudaFreeAsync(d_nruns, s); cudaFreeAsync(d_reps, s); cudaFreeAsync(d_bad, s);
    if (rc) return rc;
    h_info[0] = ok ? 1 : 0;
    return 0;
}
