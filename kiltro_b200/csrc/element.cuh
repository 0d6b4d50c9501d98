// Element-level physics, shared by the batched seam kernel (element.cu) and the fused assembly (assemble.cu).
//
// Restates, per element, what the reference evaluates with einsum/inv/det:
//   FunctionSpace.py:5-62          Q1 tables at 2 / 2x2 Gauss points, abscissa = the truncated literal 0.57735027
//   VariationalPrinciple.py:42-69  GetLocalMass
//   VariationalPrinciple.py:117-173 Temperature.GetLocalConvection + ConstitutiveConvectionIntegrand (scalar-broadcast
//                                  advection term)  -> KB_MODE_REFERENCE
//   VariationalPrinciple.py:216-261,311-342  TemperaturePG (Petrov-Galerkin weights from parent-space gradients)
//   MaterialBase.py:42-58          HeatDiffusion + HeatConvection (W (x) u.gradN)  -> KB_MODE_CONSISTENT
#pragma once
#include "common.cuh"

namespace kbel {

// Gauss abscissa and the two values the linear Lagrange functions take there (FunctionSpace.py:20-24,30,43)
constexpr double GZ = 0.57735027;
constexpr double NLO = 0.5 * (1.0 - GZ);     // N at the far Gauss point
constexpr double NHI = 0.5 * (1.0 + GZ);     // N at the near Gauss point

// 1-D: N_a(z_g): a = local node, g = Gauss point (z_0 = -GZ, z_1 = +GZ)
__device__ __forceinline__ double N1(int a, int g) { return (a == g) ? NHI : NLO; }

struct ElemOpts {
    double k, rho_cv, area, q;
    int mode;      // KB_MODE_*
    int petrov;    // Petrov-Galerkin weights
    int want_mass;
};

// Petrov-Galerkin coefficient alpha/|u| (VariationalPrinciple.py:311-330); 0 when np.isclose(Pe, 0)
__device__ __forceinline__ double pg_coefficient(double unorm, double h, double k) {
    const double Pe = unorm * h / (2.0 * k);
    if (!(fabs(Pe) > 1.0e-8)) return 0.0;           // np.isclose(Pe, 0.0): |Pe| <= atol = 1e-8 (NaN -> not close, but
                                                     // then everything is NaN anyway)
    const double alpha = 1.0 / tanh(Pe) - 1.0 / fabs(Pe);
    return alpha / unorm;
}

// ---------------------------------------------------------------------------------------------- quad-4, 2x2 Gauss
// X[a][d], U[a][d]: nodal coordinates / velocities in local node order (counter-clockwise).
// Ke[16] row-major (convectionel.ravel()), Fe[4], Me[16].
__device__ __forceinline__ void quad4(const double (&X)[4][2], const double (&U)[4][2], const ElemOpts &o,
                                      double (&Ke)[16], double (&Fe)[4], double (&Me)[16]) {
    // mean velocity, element length |X1 - X0|, PG coefficient                       VariationalPrinciple.py:122-124
    const double ub0 = (U[0][0] + U[1][0] + U[2][0] + U[3][0]) * 0.25;
    const double ub1 = (U[0][1] + U[1][1] + U[2][1] + U[3][1]) * 0.25;
    double coef = 0.0;
    if (o.petrov) {
        const double dx = X[1][0] - X[0][0], dy = X[1][1] - X[0][1];
        coef = pg_coefficient(sqrt(ub0 * ub0 + ub1 * ub1), sqrt(dx * dx + dy * dy), o.k);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) { Ke[i] = 0.0; Me[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < 4; ++i) Fe[i] = 0.0;

#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {          // eta  (2nd parent direction), outer loop   FunctionSpace.py:49
#pragma unroll
        for (int gj = 0; gj < 2; ++gj) {      // zeta (1st parent direction), inner loop   FunctionSpace.py:50
            const double Ne0 = N1(0, gi), Ne1 = N1(1, gi), Nz0 = N1(0, gj), Nz1 = N1(1, gj);
            // Bases / gBases at this Gauss point, node_arranger [0,1,3,2]             FunctionSpace.py:53-59
            const double B[4] = {Ne0 * Nz0, Ne0 * Nz1, Ne1 * Nz1, Ne1 * Nz0};
            const double gz[4] = {Ne0 * -0.5, Ne0 * 0.5, Ne1 * 0.5, Ne1 * -0.5};      // d/dzeta
            const double ge[4] = {-0.5 * Nz0, -0.5 * Nz1, 0.5 * Nz1, 0.5 * Nz0};      // d/deta
            // ParentGradientX[i][l] = sum_j gBases[i][j] X[j][l]                       VariationalPrinciple.py:139
            double J00 = 0, J01 = 0, J10 = 0, J11 = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                J00 += gz[a] * X[a][0]; J01 += gz[a] * X[a][1];
                J10 += ge[a] * X[a][0]; J11 += ge[a] * X[a][1];
            }
            const double det = J00 * J11 - J01 * J10;
            const double idet = 1.0 / det;
            const double i00 = J11 * idet, i01 = -J01 * idet, i10 = -J10 * idet, i11 = J00 * idet;   // :140
            const double dV = fabs(det);                                                // weights are 1.0   :145
            // SpatialGradient[j][l] = sum_k Jinv[j][k] gBases[k][l]                    :142
            double G0[4], G1[4], ug[4], W[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                G0[a] = i00 * gz[a] + i01 * ge[a];
                G1[a] = i10 * gz[a] + i11 * ge[a];
                ug[a] = o.rho_cv * (ub0 * G0[a] + ub1 * G1[a]);                         // advection           :152
                W[a] = B[a] + coef * (ub0 * gz[a] + ub1 * ge[a]);                       // PG weight (parent-space grads) :321
            }
            const double ak = o.area * o.k;
            if (o.mode == KB_MODE_REFERENCE) {
                double s = 0.0;                                                         // np.dot(advection, W): scalar  :170
#pragma unroll
                for (int a = 0; a < 4; ++a) s += ug[a] * W[a];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        Ke[4 * a + b] += (ak * (G0[a] * G0[b] + G1[a] * G1[b]) + s) * dV;
            } else {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        Ke[4 * a + b] += (ak * (G0[a] * G0[b] + G1[a] * G1[b]) + W[a] * ug[b]) * dV;   // MaterialBase.py:47,58
            }
            const double aq = o.area * o.q;
#pragma unroll
            for (int a = 0; a < 4; ++a) Fe[a] += aq * B[a] * dV;                        // :154,159 (Galerkin bases)
            if (o.want_mass) {
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) Me[4 * a + b] += o.rho_cv * B[a] * B[b] * dV;          // :63-67
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- line-2, 2 Gauss
__device__ __forceinline__ void line2(const double (&X)[2], const double (&U)[2], const ElemOpts &o,
                                      double (&Ke)[4], double (&Fe)[2], double (&Me)[4]) {
    const double ub = (U[0] + U[1]) * 0.5;
    double coef = 0.0;
    if (o.petrov) coef = pg_coefficient(fabs(ub), fabs(X[1] - X[0]), o.k);
#pragma unroll
    for (int i = 0; i < 4; ++i) { Ke[i] = 0.0; Me[i] = 0.0; }
    Fe[0] = Fe[1] = 0.0;
    const double g[2] = {-0.5, 0.5};                                                    // dN/dz         FunctionSpace.py:23
#pragma unroll
    for (int gp = 0; gp < 2; ++gp) {
        const double B[2] = {N1(0, gp), N1(1, gp)};
        const double J = g[0] * X[0] + g[1] * X[1];
        const double iJ = 1.0 / J;
        const double dV = fabs(J);
        double G[2], ug[2], W[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            G[a] = iJ * g[a];
            ug[a] = o.rho_cv * (ub * G[a]);
            W[a] = B[a] + coef * (ub * g[a]);
        }
        const double ak = o.area * o.k;
        const double s = ug[0] * W[0] + ug[1] * W[1];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const double adv = (o.mode == KB_MODE_REFERENCE) ? s : W[a] * ug[b];
                Ke[2 * a + b] += (ak * (G[a] * G[b]) + adv) * dV;
                if (o.want_mass) Me[2 * a + b] += o.rho_cv * B[a] * B[b] * dV;
            }
        const double aq = o.area * o.q;
        Fe[0] += aq * B[0] * dV; Fe[1] += aq * B[1] * dV;
    }
}

}  // namespace kbel
