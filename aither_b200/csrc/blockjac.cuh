// blockjac.cuh -- flux Jacobians of the block-matrix solvers (blusgs / bdplur) as inlineable
// device functions: Rusanov inviscid Jacobian, approximate thin-shear-layer viscous Jacobian, the
// reference's Gauss-Jordan inverse, and block-times-vector products.
//
// Reference: mnucci32/aither v0.10.0 include/fluxJacobian.hpp (RusanovFluxJacobian :446-483,
// InvFluxJacobian :486-562, DelprimitiveDelConservative :610-661, ApproxTSLJacobian :664-758),
// src/matrix.cpp:57-107 (MatrixInverse). A block is the flow matrix fs x fs (fs = NS + 4), row
// major, followed by the NT x NT turbulence matrix (include/fluxJacobian.hpp:62-75); the species
// diffusion entries are written for one species only (they vanish: Kronecker - Y = 0).
#pragma once
#include "physics.cuh"
#include "turbulence.cuh"

namespace aither {

template <int NS, int NT>
struct Blk {
  static constexpr int fs = NS + 4, nf = fs * fs, n = nf + NT * NT;
};

// J = 0.5 |A| dF/dU (+ 0.5 v.n |A| on the turbulence diagonal)
template <int NS, int NT>
AITHER_HD void InvFluxJacobian(const Gas &g, const double *s, const double *area, double *J) {
  using B = Blk<NS, NT>;
  constexpr int ns = NS, fs = B::fs;
#pragma unroll
  for (int q = 0; q < B::n; ++q) J[q] = 0.0;
  const double *n = area;
  const double u = s[ns], v = s[ns + 1], w = s[ns + 2];
  const double velNorm = u * n[0] + v * n[1] + w * n[2];
  const double rho = SpeciesSum<NS>(s);
  const double gamma = Gamma<NS>(g, s);
  const double gm1 = gamma - 1.0;
  const double phi = 0.5 * gm1 * (u * u + v * v + w * w);
  const double a1 = gamma * Energy<NS>(g, s) - phi;
  const double a3 = gamma - 2.0;
#define FJ(r, c) J[(r)*fs + (c)]
#pragma unroll
  for (int ii = 0; ii < ns; ++ii) {
    const double mfi = s[ii] / rho;
#pragma unroll
    for (int jj = 0; jj < ns; ++jj) FJ(ii, jj) = velNorm * ((ii == jj ? 1.0 : 0.0) - mfi);
    FJ(ii, ns + 0) = mfi * n[0];
    FJ(ii, ns + 1) = mfi * n[1];
    FJ(ii, ns + 2) = mfi * n[2];
    FJ(ns + 0, ii) = phi * n[0] - u * velNorm;
    FJ(ns + 1, ii) = phi * n[1] - v * velNorm;
    FJ(ns + 2, ii) = phi * n[2] - w * velNorm;
    FJ(ns + 3, ii) = velNorm * (phi - a1);
  }
  FJ(ns + 0, ns) = velNorm - a3 * n[0] * u;
  FJ(ns + 1, ns) = v * n[0] - gm1 * u * n[1];
  FJ(ns + 2, ns) = w * n[0] - gm1 * u * n[2];
  FJ(ns + 3, ns) = a1 * n[0] - gm1 * u * velNorm;
  FJ(ns + 0, ns + 1) = u * n[1] - gm1 * v * n[0];
  FJ(ns + 1, ns + 1) = velNorm - a3 * n[1] * v;
  FJ(ns + 2, ns + 1) = w * n[1] - gm1 * v * n[2];
  FJ(ns + 3, ns + 1) = a1 * n[1] - gm1 * v * velNorm;
  FJ(ns + 0, ns + 2) = u * n[2] - gm1 * w * n[0];
  FJ(ns + 1, ns + 2) = v * n[2] - gm1 * w * n[1];
  FJ(ns + 2, ns + 2) = velNorm - a3 * n[2] * w;
  FJ(ns + 3, ns + 2) = a1 * n[2] - gm1 * w * velNorm;
  FJ(ns + 0, ns + 3) = gm1 * n[0];
  FJ(ns + 1, ns + 3) = gm1 * n[1];
  FJ(ns + 2, ns + 3) = gm1 * n[2];
  FJ(ns + 3, ns + 3) = gamma * velNorm;
#undef FJ
  const double half = 0.5 * area[3];
#pragma unroll
  for (int q = 0; q < B::nf; ++q) J[q] *= half;
  if (NT > 0) {  // 0.5 * turbModel::InviscidConvJacobian (src/turbulence.cpp:126-136)
    const double diag = velNorm * area[3];
    J[B::nf] = 0.5 * diag;
    J[B::nf + NT * NT - 1] = 0.5 * diag;
  }
}

// dF_Ul = 0.5 (A(Ul) + lambda I), dF_Ur = 0.5 (A(Ur) - lambda I); lambda = face spectral radius
template <int NS, int NT>
AITHER_HD void RusanovFluxJacobian(const Gas &g, const double *s, const double *area, bool positive,
                                   double *J) {
  using B = Blk<NS, NT>;
  const double specRad = InvFaceSpectralRadius<NS>(s, SoS<NS>(g, s), area);
  InvFluxJacobian<NS, NT>(g, s, area, J);
#pragma unroll
  for (int r = 0; r < B::fs; ++r) {
    const double d = 1.0 * specRad;
    J[r * B::fs + r] = positive ? J[r * B::fs + r] + d : J[r * B::fs + r] - d;
  }
  if (NT > 0) {  // 0.5 * InviscidDissJacobian (src/turbulence.cpp:138-148)
    const double velNorm = s[NS] * area[0] + s[NS + 1] * area[1] + s[NS + 2] * area[2];
    const double d = 0.5 * (fabs(velNorm) * area[3]);
    J[B::nf] = positive ? J[B::nf] + d : J[B::nf] - d;
    J[B::nf + NT * NT - 1] = positive ? J[B::nf + NT * NT - 1] + d : J[B::nf + NT * NT - 1] - d;
  }
}

// TauNormal; ref: src/utility.cpp:425-437 (vg[3 r + c] = d u_c / d x_r)
AITHER_HD void TauNormalVg(const double *vg, const double *n, double mu, double mut, double *tau) {
  const double lambda = 0.0 - (2.0 / 3.0) * (mu + mut);
  const double trace = vg[0] + vg[4] + vg[8];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double mm = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) mm += (vg[3 * r + c] + vg[3 * c + r]) * n[c];
    tau[r] = lambda * trace * n[r] + (mu + mut) * mm;
  }
}

// approximate thin-shear-layer Jacobian (Dwight) wrt conserved variables, one species
template <int NS, int NT>
AITHER_HD void ApproxTslJacobian(const Gas &g, const Transport &tr, const double *s,
                                 double lamVisc, double turbVisc, double f1, const double *area,
                                 double dist, bool left, const double *vGrad, double *J) {
  static_assert(NS == 1, "species diffusion Jacobian is written for one species");
  using B = Blk<NS, NT>;
  constexpr int ns = NS, fs = B::fs;
  double A[B::nf], P[B::nf];
#pragma unroll
  for (int q = 0; q < B::nf; ++q) A[q] = P[q] = 0.0;
#pragma unroll
  for (int q = 0; q < B::n; ++q) J[q] = 0.0;
  const double t = Temperature<NS>(g, s);
  const double mu = tr.scaling * lamVisc;
  const double mut = tr.scaling * turbVisc;
  const double *n = area;
  const double u = s[ns], v = s[ns + 1], w = s[ns + 2];
  const double velNorm = u * n[0] + v * n[1] + w * n[2];
  const double rho = SpeciesSum<NS>(s);
  const double k = MixtureEffConductivity<NS>(tr, t, s);
  const double kt = mut * Mixture<NS>(g, s).cp / TurbPrandtl(tr.turbModel);
  double tauNorm[3];
  TauNormalVg(vGrad, n, mu, mut, tauNorm);
  const double fac = left ? -1.0 : 1.0;
  constexpr double third = 1.0 / 3.0;
#define FA(r, c) A[(r)*fs + (c)]
  FA(ns + 3, 0) = -(k + kt) * t / ((mu + mut) * rho) + 0.0;
  FA(ns + 0, ns) = third * n[0] * n[0] + 1.0;
  FA(ns + 1, ns) = third * n[0] * n[1];
  FA(ns + 2, ns) = third * n[0] * n[2];
  FA(ns + 3, ns) = fac * 0.5 * dist / (mu + mut) * tauNorm[0] + third * n[0] * velNorm + u;
  FA(ns + 0, ns + 1) = third * n[1] * n[0];
  FA(ns + 1, ns + 1) = third * n[1] * n[1] + 1.0;
  FA(ns + 2, ns + 1) = third * n[1] * n[2];
  FA(ns + 3, ns + 1) = fac * 0.5 * dist / (mu + mut) * tauNorm[1] + third * n[1] * velNorm + v;
  FA(ns + 0, ns + 2) = third * n[2] * n[0];
  FA(ns + 1, ns + 2) = third * n[2] * n[1];
  FA(ns + 2, ns + 2) = third * n[2] * n[2] + 1.0;
  FA(ns + 3, ns + 2) = fac * 0.5 * dist / (mu + mut) * tauNorm[2] + third * n[2] * velNorm + w;
  FA(ns + 3, ns + 3) = (k + kt) / ((mu + mut) * rho);
#undef FA
  const double scale = area[3] * (mu + mut) / dist;
#pragma unroll
  for (int q = 0; q < B::nf; ++q) A[q] *= scale;
  const double gm1 = Gamma<NS>(g, s) - 1.0;
  const double invRho = 1.0 / rho;
#define FP(r, c) P[(r)*fs + (c)]
  FP(0, 0) = 1.0;
  FP(ns + 0, 0) = -invRho * u;
  FP(ns + 1, 0) = -invRho * v;
  FP(ns + 2, 0) = -invRho * w;
  FP(ns + 3, 0) = 0.5 * gm1 * (u * u + v * v + w * w);
  FP(ns, ns) = invRho;
  FP(ns + 3, ns) = -gm1 * u;
  FP(ns + 1, ns + 1) = invRho;
  FP(ns + 3, ns + 1) = -gm1 * v;
  FP(ns + 2, ns + 2) = invRho;
  FP(ns + 3, ns + 2) = -gm1 * w;
  FP(ns + 3, ns + 3) = gm1;
#undef FP
  // MatrixMultiply (src/matrix.cpp:197-212): result(rr, ii) += L(rr, cc) R(cc, ii), cc outermost
#pragma unroll
  for (int cc = 0; cc < fs; ++cc)
#pragma unroll
    for (int rr = 0; rr < fs; ++rr)
#pragma unroll
      for (int ii = 0; ii < fs; ++ii) J[rr * fs + ii] += A[rr * fs + cc] * P[cc * fs + ii];
  if (NT > 0) {  // fac * turbModel::ViscJac (src/turbulence.cpp:485-498, :768-781)
    const double length = area[3] / dist;
    const double mt = IsSst(tr.turbModel) ? turbVisc : rho * s[NS + 4] / s[NS + 4 + (NT > 1 ? 1 : 0)];
    J[B::nf] = fac * (tr.scaling * length / rho * (lamVisc + TurbSigmaK(tr.turbModel, f1) * mt));
    J[B::nf + NT * NT - 1] =
        fac * (tr.scaling * length / rho * (lamVisc + TurbSigmaW(tr.turbModel, f1) * mt));
  }
}

// Gauss-Jordan inverse with partial pivoting in the reference's operation order
// (src/matrix.cpp:57-107); returns false for a singular matrix
template <int N>
AITHER_HD bool MatrixInverse(double *mat) {
  double I[N * N];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) I[r * N + c] = r == c ? 1.0 : 0.0;
  for (int r = 0; r < N; ++r) {
    const int cPivot = r;
    double maxVal = 0.0;
    int rPivot = 0;
    for (int ii = r; ii < N; ++ii) {
      if (fabs(mat[ii * N + cPivot]) > maxVal) {
        maxVal = fabs(mat[ii * N + cPivot]);
        rPivot = ii;
      }
    }
    if (r != rPivot) {
      for (int c = 0; c < N; ++c) {
        double tmp = mat[r * N + c];
        mat[r * N + c] = mat[rPivot * N + c];
        mat[rPivot * N + c] = tmp;
        tmp = I[r * N + c];
        I[r * N + c] = I[rPivot * N + c];
        I[rPivot * N + c] = tmp;
      }
    }
    for (int ii = 0; ii < cPivot; ++ii) {
      const double factor = mat[r * N + ii] / mat[ii * N + ii];
      for (int c = 0; c < N; ++c) {
        mat[r * N + c] = mat[r * N + c] - factor * mat[ii * N + c];
        I[r * N + c] = I[r * N + c] - factor * I[ii * N + c];
      }
    }
    if (mat[r * N + cPivot] == 0.0) return false;
    const double normFactor = 1.0 / mat[r * N + cPivot];
    for (int c = cPivot; c < N; ++c) mat[r * N + c] *= normFactor;
    for (int c = 0; c < N; ++c) I[r * N + c] *= normFactor;
  }
  for (int r = N - 2; r >= 0; --r) {
    for (int ii = N - 1; ii > r; --ii) {
      const double factor = mat[r * N + ii];
      for (int c = 0; c < N; ++c) {
        mat[r * N + c] = mat[r * N + c] - factor * mat[ii * N + c];
        I[r * N + c] = I[r * N + c] - factor * I[ii * N + c];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < N * N; ++q) mat[q] = I[q];
  return true;
}

// out = M v for a block held in registers (ArrayMultiplication, include/fluxJacobian.hpp:76-88)
template <int NS, int NT>
AITHER_HD void BlockMult(const double *M, const double *v, double *out) {
  using B = Blk<NS, NT>;
#pragma unroll
  for (int rr = 0; rr < B::fs; ++rr) {
    double acc = 0.0;
#pragma unroll
    for (int cc = 0; cc < B::fs; ++cc) acc += M[rr * B::fs + cc] * v[cc];
    out[rr] = acc;
  }
#pragma unroll
  for (int rr = 0; rr < NT; ++rr) {
    double acc = 0.0;
#pragma unroll
    for (int cc = 0; cc < NT; ++cc) acc += M[B::nf + rr * NT + cc] * v[B::fs + cc];
    out[B::fs + rr] = acc;
  }
}

}  // namespace aither
