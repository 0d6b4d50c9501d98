// K8: structured mesh generators on device.
// Replaces kiltro/MeshGeneration/Mesh.py:14-30 (Mesh.Line) and :33-58 (Mesh.Rectangle).
//
// points are bit-identical to numpy.linspace + meshgrid: linspace(a,b,n+1)[i] = i*step + a with step=(b-a)/n
// evaluated as a multiply followed by an add (no FMA contraction), last sample forced to b.
#include "common.cuh"

__device__ __forceinline__ double linspace_at(double a, double b, double step, int64_t i, int64_t n) {
    if (i == n) return b;
    return __dadd_rn(__dmul_rn((double)i, step), a);
}

__global__ void __launch_bounds__(256) k_mesh_line(double a, double b, double step, int64_t n,
                                                   double *__restrict__ points, int32_t *__restrict__ elements) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
        points[i] = linspace_at(a, b, step, i, n);
        if (i < n) {                                   // elements[e] = [e, e+1]      Mesh.py:24-25
            int2 c = make_int2((int)i, (int)i + 1);
            reinterpret_cast<int2 *>(elements)[i] = c;
        }
    }
}

// one thread per node (x fastest): points[j*(nx+1)+i] = (x_i, y_j)                    Mesh.py:37-41
// and, for i<nx, j<ny, element e=j*nx+i = [c, c+1, c+nx+2, c+nx+1], c=j*(nx+1)+i       Mesh.py:46-52
__global__ void __launch_bounds__(256) k_mesh_rect(double x0, double x1, double sx, double y0, double y1, double sy,
                                                   int64_t nx, int64_t ny, double *__restrict__ points,
                                                   int32_t *__restrict__ elements) {
    const int64_t nnode = (nx + 1) * (ny + 1);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnode; p += stride) {
        const int64_t j = p / (nx + 1), i = p - j * (nx + 1);
        double2 xy = make_double2(linspace_at(x0, x1, sx, i, nx), linspace_at(y0, y1, sy, j, ny));
        reinterpret_cast<double2 *>(points)[p] = xy;
        if (i < nx && j < ny) {
            const int c = (int)p;
            int4 el = make_int4(c, c + 1, c + (int)nx + 2, c + (int)nx + 1);
            reinterpret_cast<int4 *>(elements)[j * nx + i] = el;
        }
    }
}

// strip of a Rectangle mesh: node rows [j_first, j_first + nrows) of the global (nx+1) x (ny+1) grid, local numbering
// x-fastest, coordinates bit-identical to the global mesh (row-block partition across GPUs)
__global__ void __launch_bounds__(256) k_mesh_rect_strip(double x0, double x1, double sx, double y0, double y1, double sy,
                                                         int64_t nx, int64_t ny, int64_t j_first, int64_t nrows,
                                                         double *__restrict__ points, int32_t *__restrict__ elements) {
    const int64_t nnode = (nx + 1) * nrows;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnode; p += stride) {
        const int64_t jl = p / (nx + 1), i = p - jl * (nx + 1);
        double2 xy = make_double2(linspace_at(x0, x1, sx, i, nx), linspace_at(y0, y1, sy, j_first + jl, ny));
        reinterpret_cast<double2 *>(points)[p] = xy;
        if (i < nx && jl < nrows - 1) {
            const int c = (int)p;
            int4 el = make_int4(c, c + 1, c + (int)nx + 2, c + (int)nx + 1);
            reinterpret_cast<int4 *>(elements)[jl * nx + i] = el;
        }
    }
}

extern "C" int kb_mesh_rectangle_strip(double x0, double y0, double x1, double y1, int64_t nx, int64_t ny,
                                       int64_t j_first, int64_t nrows, double *d_points, int32_t *d_elements,
                                       void *stream) {
    if (nx < 1 || ny < 1 || j_first < 0 || nrows < 2 || j_first + nrows > ny + 1 || !d_points || !d_elements) return KB_EINVAL;
    if ((nx + 1) * nrows > INT32_MAX) return KB_ETOOBIG;
    const double sx = (x1 - x0) / (double)nx, sy = (y1 - y0) / (double)ny;
    k_mesh_rect_strip<<<kb_grid((nx + 1) * nrows, 256, 8), 256, 0, (cudaStream_t)stream>>>(
        x0, x1, sx, y0, y1, sy, nx, ny, j_first, nrows, d_points, d_elements);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_mesh_line(double left, double right, int64_t n, double *d_points, int32_t *d_elements, void *stream) {
    if (n < 1 || !d_points || !d_elements) return KB_EINVAL;
    if (n + 1 > INT32_MAX) return KB_ETOOBIG;
    const double step = (right - left) / (double)n;      // numpy linspace: step = delta/div
    k_mesh_line<<<kb_grid(n + 1, 256, 8), 256, 0, (cudaStream_t)stream>>>(left, right, step, n, d_points, d_elements);
    KB_LAUNCH_CHECK();
    return 0;
}

extern "C" int kb_mesh_rectangle(double x0, double y0, double x1, double y1, int64_t nx, int64_t ny,
                                 double *d_points, int32_t *d_elements, void *stream) {
    if (nx < 1 || ny < 1 || !d_points || !d_elements) return KB_EINVAL;
    if ((nx + 1) * (ny + 1) > INT32_MAX) return KB_ETOOBIG;
    const double sx = (x1 - x0) / (double)nx, sy = (y1 - y0) / (double)ny;
    k_mesh_rect<<<kb_grid((nx + 1) * (ny + 1), 256, 8), 256, 0, (cudaStream_t)stream>>>(
        x0, x1, sx, y0, y1, sy, nx, ny, d_points, d_elements);
    KB_LAUNCH_CHECK();
    return 0;
}

// connectivity arrives from the reference's API as int64 (Mesh.py:20,44); the kernels use int32
__global__ void __launch_bounds__(256) k_cast_i64_i32(const int64_t *__restrict__ in, int64_t n, int32_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int32_t)in[i];
}

extern "C" int kb_cast_i64_i32(const int64_t *d_in, int64_t n, int32_t *d_out, void *stream) {
    if (n < 0 || (n > 0 && (!d_in || !d_out))) return KB_EINVAL;
    if (n == 0) return 0;
    k_cast_i64_i32<<<kb_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(d_in, n, d_out);
    KB_LAUNCH_CHECK();
    return 0;
}
