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

#define KB_HD __host__ __device__ __forceinline__

namespace kbel {

// Gauss abscissa and the two values the linear Lagrange functions take there (FunctionSpace.py:20-24,30,43)
constexpr double GZ = 0.57735027;
constexpr double NLO = 0.5 * (1.0 - GZ);     // N at the far Gauss point
constexpr double NHI = 0.5 * (1.0 + GZ);     // N at the near Gauss point

// 1-D: N_a(z_g): a = local node, g = Gauss point (z_0 = -GZ, z_1 = +GZ)
KB_HD double N1(int a, int g) { return (a == g) ? NHI : NLO; }

struct ElemOpts {
    double k, rho_cv, area, q;
    int mode;      // KB_MODE_*
    int petrov;    // Petrov-Galerkin weights
    int want_mass;
};

// Petrov-Galerkin coefficient alpha/|u| (VariationalPrinciple.py:311-330); 0 when np.isclose(Pe, 0)
KB_HD double pg_coefficient(double unorm, double h, double k) {
    const double Pe = unorm * h / (2.0 * k);
    if (!(fabs(Pe) > 1.0e-8)) return 0.0;           // np.isclose(Pe, 0.0): |Pe| <= atol = 1e-8 (NaN -> not close, but
                                                     // then everything is NaN anyway)
    const double alpha = 1.0 / tanh(Pe) - 1.0 / fabs(Pe);
    return alpha / unorm;
}

// ---------------------------------------------------------------------------------------------- quad-4, 2x2 Gauss
// reciprocal of the Jacobian determinant: fp32 seed + two Newton steps (full double accuracy to ~1 ulp, 8 instructions
// instead of the ~15 of an IEEE division); falls back to the division outside the fp32 exponent range.
KB_HD double fast_rcp(double d) {
#ifdef __CUDA_ARCH__
    const double ad = fabs(d);
    if (!(ad > 1.0e-30 && ad < 1.0e30)) return 1.0 / d;
    double x = (double)__frcp_rn((float)d);
    x = x * (2.0 - d * x);
    x = x * (2.0 - d * x);
    return x;
#else
    return 1.0 / d;                    // host build (tests/test_element_host.py harness)
#endif
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
KB_HD void quad4_general(const double (&X)[4][2], const double (&U)[4][2], const ElemOpts &o,
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

// Sum-factorised form of the same integrals (the production routine).
//
// With z_a = dN_a/dzeta = sz_a h_{p(a)}[gi] and e_a = dN_a/deta = se_a h_{q(a)}[gj]  (p = 0,0,1,1; q = 0,1,1,0;
// sz = -,+,+,-; se = -,-,+,+; h_0 = (HHI, HLO), h_1 = (HLO, HHI) indexed by the Gauss index) and
// N_a = 4 h_{p(a)}[gi] h_{q(a)}[gj], every integrand is a sum of products (function of gi) x (function of gj):
//   * J^-1 = adj(J)/det, so  dV (area k) G^T G = s_g [ n11[gj] z z^T - x[gi][gj] (z e^T + e z^T) + n00[gi] e e^T ],
//     s_g = area k / |det_g|, n11 = |dX/deta|^2, n00 = |dX/dzeta|^2, x = dX/dzeta . dX/deta.  Contracting the Gauss index
//     the constant factor does not depend on first leaves 3 + 3 + 4 numbers (P, Q, Xm) from which the 10 distinct
//     entries are sums of three or four terms;
//   * dV (rho c_v u . grad N_b) = sign(det) (t0[gj] z_b + t1[gi] e_b),  t0 = rcu x dX/deta, t1 = dX/dzeta x rcu -- no
//     division at all -- and with the weights W_a = N_a + cb0 z_a + cb1 e_a the advection matrix separates the same way.
// ~215 fp64 instructions per element instead of ~450 (rounding differs from the general routine at the 1e-16 level).
// Affine elements (parallelograms: opposite edge vectors bit-equal) take a shorter form of the same algebra (~95 fp64
// instructions).  Needs det_g of one sign at the four Gauss points (any valid element, either orientation) and a product
// of the four inside the double range; otherwise returns false with Ke untouched (Fe and Me are always valid) and the caller runs
// quad4_general for Ke.
// advection coefficients shared by the general and the affine form: per weight row a,
// C_ab = sz_b al[a][pb == pa] + se_b be[a][qb == qa]   (zero when there is no advection)
KB_HD void quad4_adv_coeffs(const double (&t0)[2], const double (&t1)[2], double ub0, double ub1, double hx, double hy,
                            bool has_adv, const ElemOpts &o, double (&alS)[4], double (&alD)[4], double (&beS)[4],
                            double (&beD)[4]) {
    constexpr double HHI = 0.5 * NHI, HLO = 0.5 * NLO;
    constexpr double HH2 = HHI * HHI, LL2 = HLO * HLO, HL = HHI * HLO;
    constexpr int PA[4] = {0, 0, 1, 1}, QA[4] = {0, 1, 1, 0};
    constexpr double SZ[4] = {-1.0, 1.0, 1.0, -1.0}, SE[4] = {-1.0, -1.0, 1.0, 1.0};
    if (has_adv) {
        const double A0[2] = {HHI * t0[0] + HLO * t0[1], HLO * t0[0] + HHI * t0[1]};          // A0[q] = sum_gj h_q t0
        const double A1[2] = {HHI * t1[0] + HLO * t1[1], HLO * t1[0] + HHI * t1[1]};          // A1[p] = sum_gi h_p t1
        constexpr double H4S = 4.0 * (HH2 + LL2), H4D = 8.0 * HL;       // 4 H[p][p'] : same / different index
        if (o.petrov) {
            // weights W_a = N_a + cb0 z_a + cb1 e_a  (parent-space gradients)                    :311-330
            const double coef = pg_coefficient(sqrt(ub0 * ub0 + ub1 * ub1), sqrt(hx * hx + hy * hy), o.k);
            const double cb0 = coef * ub0, cb1 = coef * ub1;
            const double cS0 = cb0 * (t0[0] + t0[1]), cS1 = cb1 * (t1[0] + t1[1]);
            const double u0 = 0.5 * cb0, v0 = 0.5 * cb1;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double A0p = 4.0 * A0[QA[a]] + SZ[a] * cS0, A1p = 4.0 * A1[PA[a]] + SE[a] * cS1;
                const double ua = (SZ[a] * u0) * A1[PA[a]], va = (SE[a] * v0) * A0[QA[a]];
                alS[a] = (HH2 + LL2) * A0p + va; alD[a] = (2.0 * HL) * A0p + va;
                beS[a] = (HH2 + LL2) * A1p + ua; beD[a] = (2.0 * HL) * A1p + ua;
            }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                alS[a] = H4S * A0[QA[a]]; alD[a] = H4D * A0[QA[a]];
                beS[a] = H4S * A1[PA[a]]; beD[a] = H4D * A1[PA[a]];
            }
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) { alS[a] = 0.0; alD[a] = 0.0; beS[a] = 0.0; beD[a] = 0.0; }
    }
}

// Ke = (symmetric diffusion part D, upper triangle 00 01 02 03 11 12 13 22 23 33) + advection part
KB_HD void quad4_writeout(const double (&D)[10], const double (&alS)[4], const double (&alD)[4], const double (&beS)[4],
                          const double (&beD)[4], const ElemOpts &o, double (&Ke)[16]) {
    constexpr int PA[4] = {0, 0, 1, 1}, QA[4] = {0, 1, 1, 0};
    constexpr double SZ[4] = {-1.0, 1.0, 1.0, -1.0}, SE[4] = {-1.0, -1.0, 1.0, 1.0};
    constexpr int TRI[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
    if (o.mode == KB_MODE_CONSISTENT) {                                                       // MaterialBase.py:58
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b)
                Ke[4 * a + b] = D[TRI[a][b]] + (SZ[b] * (PA[b] == PA[a] ? alS[a] : alD[a]) + SE[b] * (QA[b] == QA[a] ? beS[a] : beD[a]));
    } else {                                                             // np.dot(advection, W), a scalar     :170
        double S = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) S += SZ[a] * alS[a] + SE[a] * beS[a];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) Ke[4 * a + b] = D[TRI[a][b]] + S;
    }
}

KB_HD bool quad4_fast(const double (&X)[4][2], const double (&U)[4][2], const ElemOpts &o, double (&Ke)[16],
                      double (&Fe)[4], double (&Me)[16]) {
    constexpr double HHI = 0.5 * NHI, HLO = 0.5 * NLO;
    constexpr double HH2 = HHI * HHI, LL2 = HLO * HLO, HL = HHI * HLO;

    // everything that needs the nodal velocities, first (keeps them out of the live set of the rest)      :123
    const double ub0 = (U[0][0] + U[1][0] + U[2][0] + U[3][0]) * 0.25;
    const double ub1 = (U[0][1] + U[1][1] + U[2][1] + U[3][1]) * 0.25;
    const bool has_adv = (o.rho_cv * ub0 != 0.0) || (o.rho_cv * ub1 != 0.0);

    const double e01x = X[1][0] - X[0][0], e01y = X[1][1] - X[0][1];
    const double e32x = X[2][0] - X[3][0], e32y = X[2][1] - X[3][1];
    const double e03x = X[3][0] - X[0][0], e03y = X[3][1] - X[0][1];
    const double e12x = X[2][0] - X[1][0], e12y = X[2][1] - X[1][1];
    double alS[4], alD[4], beS[4], beD[4], D[10], dV[4];
    bool ok;
    if (e01x == e32x && e01y == e32y && e03x == e12x && e03y == e12y) {
        // ---- affine element (parallelogram: everything Mesh.Rectangle generates, MeshGeneration/Mesh.py:33-58).  The
        // Jacobian is constant, so the Gauss sums collapse: the diffusion part has 6 distinct values from 3 numbers.
        const double jzx = 0.5 * e01x, jzy = 0.5 * e01y, jex = 0.5 * e03x, jey = 0.5 * e03y;
        const double det = jzx * jey - jzy * jex, ad = fabs(det);
        ok = ad > 1.0e-60 && ad < 1.0e60;
        dV[0] = dV[1] = dV[2] = dV[3] = ad;
        if (ok) {
            const double sg = det > 0.0 ? o.rho_cv : -o.rho_cv;                  // sign(det) rho c_v
            const double rb0 = sg * ub0, rb1 = sg * ub1;
            const double t0 = rb0 * jey - rb1 * jex, t1 = rb1 * jzx - rb0 * jzy;
            const double t0v[2] = {t0, t0}, t1v[2] = {t1, t1};
            quad4_adv_coeffs(t0v, t1v, ub0, ub1, e01x, e01y, has_adv, o, alS, alD, beS, beD);
            const double s = (o.area * o.k) * fast_rcp(ad);
            const double A = s * (jey * jey + jex * jex), B = s * (jzy * jzy + jzx * jzx);
            const double hm = 0.5 * (s * (jey * jzy + jex * jzx));
            constexpr double C1 = 2.0 * (HH2 + LL2), C2 = 4.0 * HL;             // 1/3 and 1/6 for the exact abscissa
            const double cab = C1 * (A + B), dab = C2 * (A + B);
            const double d00 = cab - hm, d11 = cab + hm, d01 = C2 * B - C1 * A, d03 = C2 * A - C1 * B;
            D[0] = d00; D[1] = d01; D[2] = hm - dab; D[3] = d03; D[4] = d11;
            D[5] = d03; D[6] = -dab - hm; D[7] = d00; D[8] = d01; D[9] = d11;
        }
    } else {
        // dX/dzeta at eta-index gi, dX/deta at zeta-index gj                                VariationalPrinciple.py:139
        const double Jzx[2] = {HHI * e01x + HLO * e32x, HLO * e01x + HHI * e32x};
        const double Jzy[2] = {HHI * e01y + HLO * e32y, HLO * e01y + HHI * e32y};
        const double Jex[2] = {HHI * e03x + HLO * e12x, HLO * e03x + HHI * e12x};
        const double Jey[2] = {HHI * e03y + HLO * e12y, HLO * e03y + HHI * e12y};
        double det4[4];                                   // g = 2 gi + gj
#pragma unroll
        for (int gi = 0; gi < 2; ++gi)
#pragma unroll
            for (int gj = 0; gj < 2; ++gj) det4[2 * gi + gj] = Jzx[gi] * Jey[gj] - Jzy[gi] * Jex[gj];
        const bool pos = det4[0] > 0.0 && det4[1] > 0.0 && det4[2] > 0.0 && det4[3] > 0.0;
        const bool neg = det4[0] < 0.0 && det4[1] < 0.0 && det4[2] < 0.0 && det4[3] < 0.0;
        const double p01 = det4[0] * det4[1], p23 = det4[2] * det4[3], P4 = p01 * p23;
        ok = (pos || neg) && (P4 > 1.0e-250 && P4 < 1.0e250);
#pragma unroll
        for (int g = 0; g < 4; ++g) dV[g] = fabs(det4[g]);
        if (ok) {
            // ---- advection
            const double sg = pos ? o.rho_cv : -o.rho_cv;                    // sign(det) rho c_v
            const double rb0 = sg * ub0, rb1 = sg * ub1;
            const double t0[2] = {rb0 * Jey[0] - rb1 * Jex[0], rb0 * Jey[1] - rb1 * Jex[1]};     // per gj
            const double t1[2] = {rb1 * Jzx[0] - rb0 * Jzy[0], rb1 * Jzx[1] - rb0 * Jzy[1]};     // per gi
            quad4_adv_coeffs(t0, t1, ub0, ub1, e01x, e01y, has_adv, o, alS, alD, beS, beD);
            // ---- diffusion
            // s_g = area k / |det_g| by one batched inversion
            double s[4];
            {
                const double iP = (o.area * o.k) * fast_rcp(P4);                // P4 > 0
                const double i01 = iP * p23, i23 = iP * p01;
                s[0] = fabs(i01 * det4[1]); s[1] = fabs(i01 * det4[0]); s[2] = fabs(i23 * det4[3]); s[3] = fabs(i23 * det4[2]);
            }
            const double n11[2] = {Jey[0] * Jey[0] + Jex[0] * Jex[0], Jey[1] * Jey[1] + Jex[1] * Jex[1]};
            const double n00[2] = {Jzy[0] * Jzy[0] + Jzx[0] * Jzx[0], Jzy[1] * Jzy[1] + Jzx[1] * Jzx[1]};
            const double a00[2] = {n11[0] * s[0] + n11[1] * s[1], n11[0] * s[2] + n11[1] * s[3]};      // sum over gj, per gi
            const double a11[2] = {n00[0] * s[0] + n00[1] * s[2], n00[0] * s[1] + n00[1] * s[3]};      // sum over gi, per gj
            const double P00 = HH2 * a00[0] + LL2 * a00[1], P11 = LL2 * a00[0] + HH2 * a00[1], P01 = HL * (a00[0] + a00[1]);
            const double Q00 = HH2 * a11[0] + LL2 * a11[1], Q11 = LL2 * a11[0] + HH2 * a11[1], Q01 = HL * (a11[0] + a11[1]);
            double m[4];
#pragma unroll
            for (int gi = 0; gi < 2; ++gi)
#pragma unroll
                for (int gj = 0; gj < 2; ++gj) m[2 * gi + gj] = s[2 * gi + gj] * (Jey[gj] * Jzy[gi] + Jex[gj] * Jzx[gi]);
            // Y_q[gi] = sum_gj h_q[gj] m[gi][gj];  Xm[p][q] = sum_gi h_p[gi] Y_q[gi]
            const double Y0[2] = {HHI * m[0] + HLO * m[1], HHI * m[2] + HLO * m[3]};
            const double Y1[2] = {HLO * m[0] + HHI * m[1], HLO * m[2] + HHI * m[3]};
            const double X00 = HHI * Y0[0] + HLO * Y0[1], X01 = HHI * Y1[0] + HLO * Y1[1];
            const double X10 = HLO * Y0[0] + HHI * Y0[1], X11 = HLO * Y1[0] + HHI * Y1[1];
            D[0] = (P00 + Q00) - (X00 + X00);
            D[1] = (Q01 - P00) + (X00 - X01);
            D[2] = (X01 + X10) - (P01 + Q01);
            D[3] = (P01 - Q00) + (X00 - X10);
            D[4] = (P00 + Q11) + (X01 + X01);
            D[5] = (P01 - Q11) + (X11 - X01);
            D[6] = -(P01 + Q01) - (X00 + X11);
            D[7] = (P11 + Q11) - (X11 + X11);
            D[8] = (Q01 - P11) + (X11 - X10);
            D[9] = (P11 + Q00) + (X10 + X10);
        }
    }
    if (ok) quad4_writeout(D, alS, alD, beS, beD, o, Ke);
    // ---- source vector and mass: dV-weighted sums of products of the 1-D tables                       :154,159 / :63-67
    const double aq = o.area * o.q;
    const double B0[4] = {NHI * NHI, NHI * NLO, NLO * NLO, NLO * NHI};      // Bases[:, g=0]; other points permute it
    constexpr int PERM[4][4] = {{0, 1, 2, 3}, {1, 0, 3, 2}, {3, 2, 1, 0}, {2, 3, 0, 1}};
    if (aq != 0.0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            double f = 0.0;
#pragma unroll
            for (int g = 0; g < 4; ++g) f += (aq * B0[PERM[g][a]]) * dV[g];
            Fe[a] = f;
        }
    } else {
        // 0 * dV: exact zeros unless the element is not finite
        const double z = aq * (dV[0] + dV[1] + dV[2] + dV[3]);
#pragma unroll
        for (int a = 0; a < 4; ++a) Fe[a] = z;
    }
    if (o.want_mass) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
                double mm = 0.0;
#pragma unroll
                for (int g = 0; g < 4; ++g) mm += (o.rho_cv * (B0[PERM[g][a]] * B0[PERM[g][b]])) * dV[g];
                Me[4 * a + b] = mm;
                Me[4 * b + a] = mm;
            }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) Me[i] = 0.0;
    }
    return ok;
}

// the element routine of the seam kernel (element.cu) and of the host harness
KB_HD void quad4(const double (&X)[4][2], const double (&U)[4][2], const ElemOpts &o, double (&Ke)[16],
                 double (&Fe)[4], double (&Me)[16]) {
    if (!quad4_fast(X, U, o, Ke, Fe, Me)) {
        double fe2[4], me2[16];
        quad4_general(X, U, o, Ke, fe2, me2);
    }
}

// ---------------------------------------------------------------------------------------------- line-2, 2 Gauss
KB_HD void line2(const double (&X)[2], const double (&U)[2], const ElemOpts &o,
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

// ---------------------------------------------------------------------------------------------- SURVEY 8(f) operators
// n2  Robin / Newton-cooling term, AssemblyRobinForces (FiniteElements/Assembly.py:121-175): a volume integral over every
//     element, Ke = h sum_g N (x) N dV_g (:155,163).  The source vector is reproduced as shipped: the Gauss loop adds
//     gf[counter] = h T_inf Bases[counter, :] (:157,164), so Fe[i] = h T_inf sum_g Bases[g][i] dV_g (node and Gauss
//     indices swapped -- equal to the intended sum on undistorted elements).
// n3  characteristic-Galerkin operator, GetElementalCharacteristicGalerkin (VariationalPrinciple.py:264-308):
//     Ke = sum_g (u.grad N) (x) (u.grad N) dV_g (:300,305), Fe = q area sum_g (u.grad N) dV_g (:302,306), u = mean velocity.
KB_HD void quad4_geometry(const double (&X)[4][2], double (&G0)[4][4], double (&G1)[4][4], double (&dV)[4]) {
    constexpr double HHI = 0.5 * NHI, HLO = 0.5 * NLO;
    const double e01x = X[1][0] - X[0][0], e01y = X[1][1] - X[0][1];
    const double e32x = X[2][0] - X[3][0], e32y = X[2][1] - X[3][1];
    const double e03x = X[3][0] - X[0][0], e03y = X[3][1] - X[0][1];
    const double e12x = X[2][0] - X[1][0], e12y = X[2][1] - X[1][1];
    const double Jzx[2] = {HHI * e01x + HLO * e32x, HLO * e01x + HHI * e32x};
    const double Jzy[2] = {HHI * e01y + HLO * e32y, HLO * e01y + HHI * e32y};
    const double Jex[2] = {HHI * e03x + HLO * e12x, HLO * e03x + HHI * e12x};
    const double Jey[2] = {HHI * e03y + HLO * e12y, HLO * e03y + HHI * e12y};
#pragma unroll
    for (int gi = 0; gi < 2; ++gi)
#pragma unroll
        for (int gj = 0; gj < 2; ++gj) {
            const int g = 2 * gi + gj;
            const double J00 = Jzx[gi], J01 = Jzy[gi], J10 = Jex[gj], J11 = Jey[gj];
            const double det = J00 * J11 - J01 * J10, idet = 1.0 / det;
            dV[g] = fabs(det);
            const double i00 = J11 * idet, i01 = -J01 * idet, i10 = -J10 * idet, i11 = J00 * idet;
            const double hE0 = gi == 0 ? HHI : HLO, hE1 = gi == 0 ? HLO : HHI;
            const double hZ0 = gj == 0 ? HHI : HLO, hZ1 = gj == 0 ? HLO : HHI;
            const double z[4] = {-hE0, hE0, hE1, -hE1}, e[4] = {-hZ0, -hZ1, hZ1, hZ0};
#pragma unroll
            for (int a = 0; a < 4; ++a) { G0[g][a] = i00 * z[a] + i01 * e[a]; G1[g][a] = i10 * z[a] + i11 * e[a]; }
        }
}

// Bases[a][g] of the quad (FunctionSpace.py:53-59): Bases[:, 0] = B0, the other Gauss points permute it
KB_HD double quad4_basis(int a, int g) {
    const double B0[4] = {NHI * NHI, NHI * NLO, NLO * NLO, NLO * NHI};
    constexpr int PERM[4][4] = {{0, 1, 2, 3}, {1, 0, 3, 2}, {3, 2, 1, 0}, {2, 3, 0, 1}};
    return B0[PERM[g][a]];
}

KB_HD void quad4_robin(const double (&X)[4][2], double h, double t_inf, double (&Ke)[16], double (&Fe)[4]) {
    double G0[4][4], G1[4][4], dV[4];
    quad4_geometry(X, G0, G1, dV);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            double m = 0.0;
#pragma unroll
            for (int g = 0; g < 4; ++g) m += (h * (quad4_basis(a, g) * quad4_basis(b, g))) * dV[g];
            Ke[4 * a + b] = m;
        }
        double f = 0.0;
#pragma unroll
        for (int g = 0; g < 4; ++g) f += (h * t_inf * quad4_basis(g, a)) * dV[g];     // as shipped: Bases[g][a]
        Fe[a] = f;
    }
}

KB_HD void quad4_chargalerkin(const double (&X)[4][2], const double (&U)[4][2], double qa, double (&Ke)[16],
                              double (&Fe)[4]) {
    double G0[4][4], G1[4][4], dV[4];
    quad4_geometry(X, G0, G1, dV);
    const double ub0 = (U[0][0] + U[1][0] + U[2][0] + U[3][0]) * 0.25;
    const double ub1 = (U[0][1] + U[1][1] + U[2][1] + U[3][1]) * 0.25;
#pragma unroll
    for (int i = 0; i < 16; ++i) Ke[i] = 0.0;
#pragma unroll
    for (int a = 0; a < 4; ++a) Fe[a] = 0.0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        double ug[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) ug[a] = ub0 * G0[g][a] + ub1 * G1[g][a];            // USpatialGradient   :296
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) Ke[4 * a + b] += (ug[a] * ug[b]) * dV[g];
            Fe[a] += (qa * ug[a]) * dV[g];
        }
    }
}

KB_HD void line2_robin(const double (&X)[2], double h, double t_inf, double (&Ke)[4], double (&Fe)[2]) {
    const double dV = fabs(0.5 * (X[1] - X[0]));                                          // both Gauss points
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int b = 0; b < 2; ++b) Ke[2 * a + b] = (h * (N1(a, 0) * N1(b, 0))) * dV + (h * (N1(a, 1) * N1(b, 1))) * dV;
        Fe[a] = (h * t_inf * N1(0, a)) * dV + (h * t_inf * N1(1, a)) * dV;
    }
}

KB_HD void line2_chargalerkin(const double (&X)[2], const double (&U)[2], double qa, double (&Ke)[4], double (&Fe)[2]) {
    const double J = -0.5 * X[0] + 0.5 * X[1], iJ = 1.0 / J, dV = fabs(J);
    const double ub = (U[0] + U[1]) * 0.5;
    const double ug[2] = {ub * (-0.5 * iJ), ub * (0.5 * iJ)};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int b = 0; b < 2; ++b) Ke[2 * a + b] = (ug[a] * ug[b]) * dV + (ug[a] * ug[b]) * dV;
        Fe[a] = (qa * ug[a]) * dV + (qa * ug[a]) * dV;
    }
}

}  // namespace kbel
