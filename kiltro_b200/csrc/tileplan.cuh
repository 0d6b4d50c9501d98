// Tile plan of the fused assembly (shared between tileplan.cu and assemble_fused.cu).
//
// A tile = the set of CSR rows (nodes) one CTA owns + the elements incident to them.  For meshes made by
// Mesh.Rectangle (node numbering x-fastest, known nx_nodes x ny_nodes) tiles are TX x TY patches of nodes, which
// keeps the redundant element work at the patch rim to (TX+1)(TY+1)/(TX TY) - 1 = 13.8 %; for arbitrary connectivity
// (and for line meshes) tiles are runs of R consecutive rows -- always correct, more rim work on 2-D meshes.
//
// Plan arrays (all built on device by kb_tile_plan_build, consumed by k_assemble_fused):
//   tile_row_ptr[ntiles+1], tile_rows[nnode]      rows of tile t, in tile order
//   tile_elem_ptr[ntiles+1], tile_elems[...]      distinct incident elements of tile t, ascending
//   pair_ptr[nnode+1] (indexed by position in tile_rows), pair_rec[...]   for every (row, incident element), ascending
//        element: bits 18..30 slot of the element in its tile, bits 16..17 local node a of the row in the element,
//        bits 4b..4b+3 position (inside the row) of column elements[e][b]; bit 31 set when a position is >= 16
//        (the kernel then searches the row's column list instead).
#pragma once
#include "common.cuh"

namespace kbtile {

constexpr int THREADS = 256;
constexpr int EMAX = 256;        // elements per tile (one per thread in the element phase)
constexpr int RMAX = 256;        // rows per tile
constexpr int VMAX = 2560;       // CSR entries per tile staged in shared memory
constexpr int TX = 15, TY = 15;  // node patch of structured tiles -> 16 x 16 = 256 incident elements
constexpr int RCHUNK_2D = 96;    // rows per tile, unstructured quad meshes
constexpr int RCHUNK_1D = 224;   // rows per tile, line meshes

struct Geom {
    int structured, nxn, nyn, ntx, nty, R, ntiles;
    int64_t nnode;
};

static inline Geom make_geom(int64_t nnode, int npe, int64_t nxn, int64_t nyn) {
    Geom g;
    g.nnode = nnode;
    g.structured = (npe == 4 && nxn > 1 && nyn > 1 && nxn * nyn == nnode) ? 1 : 0;
    g.nxn = (int)nxn; g.nyn = (int)nyn;
    g.R = npe == 4 ? RCHUNK_2D : RCHUNK_1D;
    if (g.structured) {
        g.ntx = (int)((nxn + TX - 1) / TX); g.nty = (int)((nyn + TY - 1) / TY);
        g.ntiles = g.ntx * g.nty;
    } else {
        g.ntx = g.nty = 0;
        g.ntiles = (int)((nnode + g.R - 1) / g.R);
    }
    return g;
}

__host__ __device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }

// tile of row i and the position of i in the tile-ordered row list
__host__ __device__ __forceinline__ void tile_of(const Geom &g, int i, int &t, int &slot) {
    if (g.structured) {
        const int iy = i / g.nxn, ix = i - iy * g.nxn;
        const int ty = iy / TY, tx = ix / TX;
        const int th = imin(TY, g.nyn - ty * TY), tw = imin(TX, g.nxn - tx * TX);
        t = ty * g.ntx + tx;
        slot = ty * TY * g.nxn + th * (tx * TX) + (iy - ty * TY) * tw + (ix - tx * TX);
    } else {
        t = i / g.R;
        slot = i;
    }
}

__host__ __device__ __forceinline__ int tile_row_start(const Geom &g, int t) {
    if (t >= g.ntiles) return (int)g.nnode;
    if (g.structured) {
        const int ty = t / g.ntx, tx = t - ty * g.ntx;
        const int th = imin(TY, g.nyn - ty * TY);
        return ty * TY * g.nxn + th * (tx * TX);
    }
    return t * g.R;
}


// ---- tile templates (tiletmpl.cu builds them, assemble_tmpl.cu consumes them)
// On meshes whose numbering repeats (everything Mesh.Rectangle makes, row-block strips of it, ...) almost all tiles have
// the same LOCAL structure: same rows-in-runs layout, same element slots, same (entry <- element-matrix components)
// lists.  The builder hashes each tile's local plan, keeps one image per distinct plan (a "template") and a class id
// per tile; the kernel keeps the current template in shared memory, so per tile only the run bases come from HBM.
// A template lists, for every CSR entry of the tile (tile order), up to 4 sources as BYTE offsets into the kernel's
// element-matrix staging area (component-major, KSTRIDE doubles per component; unused sources point at a zero word), in
// ascending element order = the order of the reference's sequential `+=` (Assembly.py:46,60-61).
constexpr int RUNMAX = 31;                 // runs of consecutive rows per tile
constexpr int CMAX = 64;                   // distinct templates per mesh
constexpr int KSTRIDE = EMAX + 1;          // doubles per component row (odd: spreads components over the banks)
constexpr int KZERO = 16 * KSTRIDE;        // the zero word behind the 16 components
constexpr int FSTRIDE = EMAX + 1;
constexpr int FZERO = 4 * FSTRIDE;

struct TileTmpl {
    int nent, nrow, nrun, ne;
    uint16_t run_ent[RUNMAX + 1];          // local entry offset of each run (+ total)
    uint16_t run_row[RUNMAX + 1];          // local row offset of each run (+ nrow)
    uint16_t ent_src[VMAX][4];             // byte offsets into the Ke staging
    uint16_t row_src[RMAX][4];             // byte offsets into the Fe staging
    uint8_t ent_run[VMAX];
    uint8_t row_run[RMAX];
};
static_assert(sizeof(TileTmpl) % 16 == 0, "templates are copied with 16-byte words");

}  // namespace kbtile
