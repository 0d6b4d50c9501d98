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
// reciprocal of the Jacobian determinant: fp32 seed + two Newton steps (full double accuracy to ~1 ulp, 8 instructions
// instead of the ~15 of an IEEE division); falls back to the division outside the fp32 exponent range.
__device__ __forceinline__ double fast_rcp(double d) {
    const double ad = fabs(d);
    if (!(ad > 1.0e-30 && ad < 1.0e30)) return 1.0 / d;
    double x = (double)__frcp_rn((float)d);
    x = x * (2.0 - d * x);
    x = x * (2.0 - d * x);
    return x;
}

// X[a][d], U[a][d]: nodal coordinates / velocities in local node order (counter-clockwise).
// Ke[16] row-major (convectionel.ravel()), Fe[4], Me[16].
//
// Same integrals as VariationalPrinciple.py:117-173 / :216-261 / :42-69, organised for the FP64 pipe:
//   * the parent-space Jacobian of a Q1 quad is affine in (zeta, eta): dX/dzeta depends on eta only and is a blend of the
//     two zeta-edges, dX/deta depends on zeta only -> 2+2 edge blends instead of four 4-term sums per Gauss point;
//   * the geometry of the four Gauss points is evaluated first (four independent chains: determinant, reciprocal,
//     inverse, gradients), then the accumulation streams over them;
//   * the diffusion part is symmetric: 10 entries are accumulated and mirrored at the end;
//   * dV (and area*k*dV) are folded into one operand of each product before the 4x4 update.
// Rounding differs from NumPy's einsum order at the 1e-16 level (tests: <= 1e-12 relative to the reference's output).
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
    const double rb0 = o.rho_cv * ub0, rb1 = o.rho_cv * ub1;         // rho c_v u   (advection, :152)
    const double cb0 = coef * ub0, cb1 = coef * ub1;                 // alpha/|u| u (PG weights, :321)
    const double ak = o.area * o.k;

    // edges: dX/dzeta = Ne0/2 (X1-X0) + Ne1/2 (X2-X3) ; dX/deta = Nz0/2 (X3-X0) + Nz1/2 (X2-X1)
    const double e01x = X[1][0] - X[0][0], e01y = X[1][1] - X[0][1];
    const double e32x = X[2][0] - X[3][0], e32y = X[2][1] - X[3][1];
    const double e03x = X[3][0] - X[0][0], e03y = X[3][1] - X[0][1];
    const double e12x = X[2][0] - X[1][0], e12y = X[2][1] - X[1][1];
    constexpr double HHI = 0.5 * NHI, HLO = 0.5 * NLO;               // |dN/dxi| at a Gauss point: N/2
    // index 0: Gauss coordinate -GZ, 1: +GZ
    const double Jzx[2] = {HHI * e01x + HLO * e32x, HLO * e01x + HHI * e32x};
    const double Jzy[2] = {HHI * e01y + HLO * e32y, HLO * e01y + HHI * e32y};
    const double Jex[2] = {HHI * e03x + HLO * e12x, HLO * e03x + HHI * e12x};
    const double Jey[2] = {HHI * e03y + HLO * e12y, HLO * e03y + HHI * e12y};

    // the four Jacobian determinants and their reciprocals by one batched inversion (one reciprocal + 9 products
    // instead of four reciprocals; falls back to plain divisions when the product leaves the safe exponent range)
    double det4[4], idet4[4];
#pragma unroll
    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
        for (int gj = 0; gj < 2; ++gj) det4[2 * gi + gj] = Jzx[gi] * Jey[gj] - Jzy[gi] * Jex[gj];
    {
        const double p01 = det4[0] * det4[1], p23 = det4[2] * det4[3], P = p01 * p23, aP = fabs(P);
        if (aP > 1.0e-250 && aP < 1.0e250) {
            const double iP = fast_rcp(P);
            const double i01 = iP * p23, i23 = iP * p01;
            idet4[0] = i01 * det4[1]; idet4[1] = i01 * det4[0]; idet4[2] = i23 * det4[3]; idet4[3] = i23 * det4[2];
        } else {
#pragma unroll
            for (int g = 0; g < 4; ++g) idet4[g] = 1.0 / det4[g];
        }
    }
    double D[10];                 // symmetric diffusion part, upper triangle (00 01 02 03 11 12 13 22 23 33)
    double C[16];                 // advection part
    double dV[4];
#pragma unroll
    for (int i = 0; i < 10; ++i) D[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) C[i] = 0.0;
    double S = 0.0;               // reference mode: the broadcast scalar
    const bool consistent = o.mode == KB_MODE_CONSISTENT;
    const bool has_adv = (rb0 != 0.0) || (rb1 != 0.0);
    // Gauss points g = 2*gi + gj (gi: eta index, outer; gj: zeta index, inner)               FunctionSpace.py:49-50
#pragma unroll
    for (int gi = 0; gi < 2; ++gi) {
#pragma unroll
        for (int gj = 0; gj < 2; ++gj) {
            const int g = 2 * gi + gj;
            const double J00 = Jzx[gi], J01 = Jzy[gi], J10 = Jex[gj], J11 = Jey[gj];   // ParentGradientX          :139
            const double idet = idet4[g];
            dV[g] = fabs(det4[g]);                                                      // weights are 1.0          :145
            const double i00 = J11 * idet, i01 = -J01 * idet, i10 = -J10 * idet, i11 = J00 * idet;   // inverse   :140
            // gBases at this point: d/dzeta = (-hE0, hE0, hE1, -hE1), d/deta = (-hZ0, -hZ1, hZ1, hZ0)
            const double hE0 = gi == 0 ? HHI : HLO, hE1 = gi == 0 ? HLO : HHI;
            const double hZ0 = gj == 0 ? HHI : HLO, hZ1 = gj == 0 ? HLO : HHI;
            const double p0 = i00 * hE0, p1 = i00 * hE1, q0 = i01 * hZ0, q1 = i01 * hZ1;
            const double G0[4] = {-p0 - q0, p0 - q1, p1 + q1, q0 - p1};                 // dN/dx                    :142
            const double r0 = i10 * hE0, r1 = i10 * hE1, s0 = i11 * hZ0, s1 = i11 * hZ1;
            const double G1[4] = {-r0 - s0, r0 - s1, r1 + s1, s0 - r1};                 // dN/dy
            const double w = ak * dV[g];
            double H0[4], H1[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { H0[a] = w * G0[a]; H1[a] = w * G1[a]; }
            int t = 0;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = a; b < 4; ++b) { D[t] += H0[a] * G0[b] + H1[a] * G1[b]; ++t; }   // (area k G)^T G dV     :150,169
            if (has_adv) {
                double ug[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) ug[a] = (rb0 * G0[a] + rb1 * G1[a]) * dV[g];          // advection dV     :152
                // weights: Bases (+ PG correction from the parent-space gradients)        FunctionSpace.py:53-59, :321
                const double Ne0 = 2.0 * hE0, Ne1 = 2.0 * hE1, Nz0 = 2.0 * hZ0, Nz1 = 2.0 * hZ1;
                const double W[4] = {Ne0 * Nz0 + (cb0 * -hE0 + cb1 * -hZ0), Ne0 * Nz1 + (cb0 * hE0 + cb1 * -hZ1),
                                     Ne1 * Nz1 + (cb0 * hE1 + cb1 * hZ1), Ne1 * Nz0 + (cb0 * -hE1 + cb1 * hZ0)};
                if (consistent) {
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int b = 0; b < 4; ++b) C[4 * a + b] += W[a] * ug[b];                       // MaterialBase.py:58
                } else {
#pragma unroll
                    for (int a = 0; a < 4; ++a) S += ug[a] * W[a];                                      // np.dot(advection, W) :170
                }
            }
        }
    }
    // ---- write out: mirror the diffusion part, add the advection part
    {
        int t = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
                Ke[4 * a + b] = D[t] + (consistent ? C[4 * a + b] : S);
                if (b != a) Ke[4 * b + a] = D[t] + (consistent ? C[4 * b + a] : S);
                ++t;
            }
    }
    // source vector: area q N dV (Galerkin bases)                                                        :154,159
    const double aq = o.area * o.q;
    const double B0[4] = {NHI * NHI, NHI * NLO, NLO * NLO, NLO * NHI};      // Bases[:, g=0]; other points permute it
    // Bases[a][g]: g=0 (B0[0],B0[1],B0[2],B0[3]); g=1: zeta flips -> (B0[1],B0[0],B0[3],B0[2]);
    //              g=2: eta flips -> (B0[3],B0[2],B0[1],B0[0]); g=3: both -> (B0[2],B0[3],B0[0],B0[1])
    constexpr int PERM[4][4] = {{0, 1, 2, 3}, {1, 0, 3, 2}, {3, 2, 1, 0}, {2, 3, 0, 1}};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        double f = 0.0;
#pragma unroll
        for (int g = 0; g < 4; ++g) f += (aq * B0[PERM[g][a]]) * dV[g];
        Fe[a] = f;
    }
    if (o.want_mass) {                                                                                   // :63-67
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
                double m = 0.0;
#pragma unroll
                for (int g = 0; g < 4; ++g) m += (o.rho_cv * (B0[PERM[g][a]] * B0[PERM[g][b]])) * dV[g];
                Me[4 * a + b] = m;
                Me[4 * b + a] = m;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) Me[i] = 0.0;
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
