// Host build of the element math of the CUDA kernels (kiltro_b200/csrc/element.cuh is __host__ __device__): lets the
// CPU-only test tier check the exact formulas the kernels run against the reference-generated golden element matrices.
// Test infrastructure only -- nothing in the product links this.
#include "../../kiltro_b200/csrc/element.cuh"

using namespace kbel;

static ElemOpts make_opts(const double *mat, int mode, int petrov, int want_mass) {
    ElemOpts o;                                     // mat = (k, rho, c_v, area, q) like kb_material
    o.k = mat[0]; o.rho_cv = mat[1] * mat[2]; o.area = mat[3]; o.q = mat[4];
    o.mode = mode; o.petrov = petrov; o.want_mass = want_mass;
    return o;
}

// which: 0 = quad4 (sum-factorised, production), 1 = quad4_general
extern "C" int host_quad4(int which, const double *points, const int *elements, long nelem, const double *velocity,
                          const double *mat, int mode, int petrov, int want_mass, double *Ke, double *Fe, double *Me) {
    const ElemOpts o = make_opts(mat, mode, petrov, want_mass);
    for (long e = 0; e < nelem; ++e) {
        double X[4][2], U[4][2], ke[16], fe[4], me[16];
        for (int a = 0; a < 4; ++a) {
            const int n = elements[4 * e + a];
            X[a][0] = points[2 * n]; X[a][1] = points[2 * n + 1];
            U[a][0] = velocity ? velocity[2 * n] : 0.0; U[a][1] = velocity ? velocity[2 * n + 1] : 0.0;
        }
        if (which == 0) quad4(X, U, o, ke, fe, me); else quad4_general(X, U, o, ke, fe, me);
        for (int i = 0; i < 16; ++i) { Ke[16 * e + i] = ke[i]; Me[16 * e + i] = me[i]; }
        for (int i = 0; i < 4; ++i) Fe[4 * e + i] = fe[i];
    }
    return 0;
}

extern "C" int host_line2(const double *points, const int *elements, long nelem, const double *velocity,
                          const double *mat, int mode, int petrov, int want_mass, double *Ke, double *Fe, double *Me) {
    const ElemOpts o = make_opts(mat, mode, petrov, want_mass);
    for (long e = 0; e < nelem; ++e) {
        const int n0 = elements[2 * e], n1 = elements[2 * e + 1];
        double X[2] = {points[n0], points[n1]}, U[2] = {velocity ? velocity[n0] : 0.0, velocity ? velocity[n1] : 0.0};
        double ke[4], fe[2], me[4];
        line2(X, U, o, ke, fe, me);
        for (int i = 0; i < 4; ++i) { Ke[4 * e + i] = ke[i]; Me[4 * e + i] = me[i]; }
        for (int i = 0; i < 2; ++i) Fe[2 * e + i] = fe[i];
    }
    return 0;
}

// op: 1 = Robin (p0 = h, p1 = T_inf), 2 = characteristic-Galerkin (p0 = q * area); npe = 4 (quad) or 2 (line)
extern "C" int host_operator(int npe, int op, const double *points, const int *elements, long nelem, const double *velocity,
                             double p0, double p1, double *Ke, double *Fe) {
    for (long e = 0; e < nelem; ++e) {
        if (npe == 4) {
            double X[4][2], U[4][2], ke[16], fe[4];
            for (int a = 0; a < 4; ++a) {
                const int n = elements[4 * e + a];
                X[a][0] = points[2 * n]; X[a][1] = points[2 * n + 1];
                U[a][0] = velocity ? velocity[2 * n] : 0.0; U[a][1] = velocity ? velocity[2 * n + 1] : 0.0;
            }
            if (op == 1) quad4_robin(X, p0, p1, ke, fe); else quad4_chargalerkin(X, U, p0, ke, fe);
            for (int i = 0; i < 16; ++i) Ke[16 * e + i] = ke[i];
            for (int i = 0; i < 4; ++i) Fe[4 * e + i] = fe[i];
        } else {
            const int n0 = elements[2 * e], n1 = elements[2 * e + 1];
            double X[2] = {points[n0], points[n1]}, U[2] = {velocity ? velocity[n0] : 0.0, velocity ? velocity[n1] : 0.0};
            double ke[4], fe[2];
            if (op == 1) line2_robin(X, p0, p1, ke, fe); else line2_chargalerkin(X, U, p0, ke, fe);
            for (int i = 0; i < 4; ++i) Ke[4 * e + i] = ke[i];
            for (int i = 0; i < 2; ++i) Fe[2 * e + i] = fe[i];
        }
    }
    return 0;
}
