// turbulence.cuh -- point functions of the two-equation RANS models (k-omega Wilcox 2006 and
// Menter SST 2003) as inlineable device functions: eddy viscosity and blending at a face, the
// diffusion coefficients, the source terms and the turbulence-equation spectral radii.
//
// Reference: mnucci32/aither v0.10.0 src/turbulence.cpp (cited per function), constants
// include/turbulence.hpp:391-398 (Wilcox) and :489-501 (SST). The model is a run-time switch on
// `turbModel` (two short branches), the equation counts are template parameters.
// Tensor convention as the reference's velocityGrad_: vg[3 r + c] = d u_c / d x_r.
#pragma once
#include "physics.cuh"

namespace aither {

namespace kw {
constexpr double gamma = 0.52, betaStar = 0.09, sigma = 0.5, sigmaStar = 0.6, sigmaD0 = 0.125,
                 beta0 = 0.0708, clim = 0.875;
}
namespace sst {
constexpr double betaStar = 0.09, sigmaK1 = 0.85, sigmaK2 = 1.0, sigmaW1 = 0.5, sigmaW2 = 0.856,
                 beta1 = 0.075, beta2 = 0.0828, gamma1 = 5.0 / 9.0, gamma2 = 0.44, a1 = 0.31,
                 kProd2Dest = 10.0;
}

AITHER_HD bool IsSst(int turbModel) { return turbModel == AITHER_TURB_SST; }
// ref: src/turbulence.cpp:592-595
AITHER_HD double Blended(double c1, double c2, double f1) { return f1 * c1 + (1.0 - f1) * c2; }
// ref: include/turbulence.hpp:476-477 (Wilcox), :599-604 (SST)
AITHER_HD double TurbSigmaK(int turbModel, double f1) {
  return IsSst(turbModel) ? Blended(sst::sigmaK1, sst::sigmaK2, f1) : kw::sigmaStar;
}
AITHER_HD double TurbSigmaW(int turbModel, double f1) {
  return IsSst(turbModel) ? Blended(sst::sigmaW1, sst::sigmaW2, f1) : kw::sigma;
}
// ref: include/turbulence.hpp:463, :577
AITHER_HD double TurbWallBeta(int turbModel) { return IsSst(turbModel) ? sst::beta1 : kw::beta0; }

AITHER_HD double DDotTrans(const double *a, const double *b) {
  // tensor::DoubleDotTrans: sum of the elementwise product; include/tensor.hpp:353-356
  double sum = 0.0;
#pragma unroll
  for (int q = 0; q < 9; ++q) sum += a[q] * b[q];
  return sum;
}
AITHER_HD double Dot3(const double *a, const double *b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

// SST cross diffusion CD_kw; ref: include/turbulence.hpp:528-537
AITHER_HD double SstCdkw(double rho, double omega, const double *kg, const double *wg) {
  return fmax(2.0 * rho * sst::sigmaW2 / omega * Dot3(kg, wg), 1.0e-10);
}

// turbModel::EddyViscAndBlending at a face; ref: src/turbulence.cpp:405-423 (Wilcox, OmegaTilda
// :329-342), :663-684 (SST: Alpha1-3 :597-615, F1/F2 :582-590, EddyVisc :570-580)
AITHER_HD void EddyViscAndBlending(int turbModel, double scaling, double rho, double tke,
                                   double omega, const double *vg, const double *kg,
                                   const double *wg, double mu, double wallDist, double *mut,
                                   double *f1, double *f2) {
  const double trace = vg[0] + vg[4] + vg[8];
  if (!IsSst(turbModel)) {
    *f1 = 1.0;
    *f2 = 0.0;
    double ss = 0.0;  // sHat : sHat, sHat = sym(G) - tr(G)/3 I
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double sh = 0.5 * (vg[3 * r + c] + vg[3 * c + r]) - 1.0 / 3.0 * trace * (r == c ? 1.0 : 0.0);
        ss += sh * sh;
      }
    const double omegaTilda = fmax(omega, scaling * kw::clim * sqrt(2.0 * ss / kw::betaStar));
    *mut = rho * tke / omegaTilda;
    return;
  }
  const double dE = wallDist + kEps;
  const double alpha1 = scaling * sqrt(tke) / (sst::betaStar * omega * dE);
  const double alpha2 = scaling * scaling * 500.0 * mu / (dE * dE * rho * omega);
  const double cdkw = SstCdkw(rho, omega, kg, wg);
  const double alpha3 = 4.0 * rho * sst::sigmaW2 * tke / (cdkw * dE * dE);
  const double arg1 = fmin(fmax(alpha1, alpha2), alpha3);
  const double a12 = arg1 * arg1;
  *f1 = tanh(a12 * a12);
  const double arg2 = fmax(2.0 * alpha1, alpha2);
  *f2 = tanh(arg2 * arg2);
  double ss = 0.0;  // S : S, S = sym(G)
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double sr = 0.5 * (vg[3 * r + c] + vg[3 * c + r]);
      ss += sr * sr;
    }
  const double meanStrainRate = sqrt(2.0 * ss);
  *mut = rho * sst::a1 * tke / fmax(sst::a1 * omega, scaling * meanStrainRate * (*f2));
}

// viscous part of the turbulence-equation spectral radius without the geometric length:
// scaling / rho (mu + sigma_k mu_t*), mu_t* = rho k / omega for Wilcox (unlimited), mu_t for SST;
// times |A|^2 / V (cell) or |A| / dist (face). ref: src/turbulence.cpp:500-527, :783-808
AITHER_HD double TurbViscSpecFactor(int turbModel, double scaling, double rho, double tke,
                                    double omega, double mu, double mut, double f1) {
  const double mt = IsSst(turbModel) ? mut : rho * tke / omega;
  return scaling / rho * (mu + TurbSigmaK(turbModel, f1) * mt);
}

// turbModel::SrcSpecRad; ref: src/turbulence.cpp:438-443, :699-704
AITHER_HD double TurbSrcSpecRad(double scaling, double omega, double vol) {
  return -2.0 * kw::betaStar * omega * vol * (1.0 / scaling);
}

// turbModel::CalcTurbSrc: {k source, omega source}; ref: src/turbulence.cpp:344-384 (Wilcox; Beta,
// FBeta, Xw :291-319), :617-661 (SST); BoussinesqReynoldsStress :55-70
AITHER_HD void TurbSource(int turbModel, double scaling, double rho, double tke, double omega,
                          const double *vg, const double *kg, const double *wg, double mut,
                          double f1, double *src, double *betaOut = nullptr) {
  const double invScaling = 1.0 / scaling;
  const double trace = vg[0] + vg[4] + vg[8];
  const double lambda = 0.0 - (2.0 / 3.0) * mut;
  double prod = 0.0;  // tau : G
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double I = r == c ? 1.0 : 0.0;
      const double tau = lambda * trace * I + mut * (vg[3 * r + c] + vg[3 * c + r]) -
                         2.0 / 3.0 * rho * tke * I;
      prod += tau * vg[3 * r + c];
    }
  const double prodRaw = scaling * prod;
  if (!IsSst(turbModel)) {
    const double tkeDest = invScaling * kw::betaStar * (rho * tke * omega);
    double vort[9], ski[9], vv[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        vort[3 * r + c] = 0.5 * (vg[3 * r + c] - vg[3 * c + r]);
        ski[3 * r + c] = 0.5 * (vg[3 * r + c] + vg[3 * c + r] - trace * (r == c ? 1.0 : 0.0));
        vv[3 * r + c] = 0.0;
      }
#pragma unroll
    for (int cc = 0; cc < 3; ++cc)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr)
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) vv[3 * rr + ii] += vort[3 * rr + cc] * vort[3 * cc + ii];
    const double bw = kw::betaStar * omega;
    const double xw = fabs(DDotTrans(vv, ski) / (bw * bw * bw)) * (scaling * scaling * scaling);
    const double beta = kw::beta0 * ((1.0 + 85.0 * xw) / (1.0 + 100.0 * xw));
    if (betaOut) *betaOut = beta;
    const double omgDest = invScaling * beta * (rho * omega * omega);
    const double tkeProd = fmax(prodRaw, 0.0);
    const double omgProd = fmax(kw::gamma * omega / tke * tkeProd, 0.0);
    const double kwDot = Dot3(kg, wg);
    const double sigmaD = kwDot <= 0.0 ? 0.0 : kw::sigmaD0;
    const double omgCd = scaling * sigmaD * (rho / omega * kwDot);
    src[0] = tkeProd - tkeDest;
    src[1] = omgProd - omgDest + omgCd;
    return;
  }
  const double cdkw = SstCdkw(rho, omega, kg, wg);
  const double gamma = Blended(sst::gamma1, sst::gamma2, f1);
  const double beta = Blended(sst::beta1, sst::beta2, f1);
  if (betaOut) *betaOut = beta;
  const double tkeDest = invScaling * sst::betaStar * (rho * tke * omega);
  const double omgDest = invScaling * beta * (rho * omega * omega);
  const double tkeProd = fmax(fmin(prodRaw, sst::kProd2Dest * tkeDest), 0.0);
  const double omgProd = fmax(gamma * rho / mut * tkeProd, 0.0);
  const double omgCd = scaling * (1.0 - f1) * cdkw;
  src[0] = tkeProd - tkeDest;
  src[1] = omgProd - omgDest + omgCd;
}

}  // namespace aither
