/* aither_oracle.c -- plain-C CPU restatement of the reference hot path.
 * TEST INFRASTRUCTURE ONLY; see aither_oracle.h for the rules and the pinning.
 *
 * Sequential, array-of-structs, the reference's own loop and accumulation
 * order. "ref:" comments cite mnucci32/aither v0.10.0 file:line.
 */
#include "aither_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EPS 1.0e-30 /* ref: include/macros.hpp.in:21 */
#define MAXEQ (AITHER_MAX_SPECIES + 6)
#define MAXFS (AITHER_MAX_SPECIES + 4)
#define MAXJAC (MAXFS * MAXFS + 4)

typedef struct {
  int ni, nj, nk, g;
  int NI, NJ, NK; /* padded extents */
  int parentBlock;
  int nsurf;
  aither_surface *surf;
  double *state;     /* padded, neq */
  double *residual;  /* ni nj nk neq */
  double *specRad;   /* ni nj nk 2 */
  double *dt;        /* ni nj nk */
  double *a, *ainv;  /* ni nj nk asz */
  double *x;         /* padded, neq */
  double *xold;      /* padded, neq (scratch for DPLUR) */
  double *consN;     /* ni nj nk neq */
  double *consNm1;   /* ni nj nk neq */
  double *mresid;    /* ni nj nk neq: matrix residual */
  double *temperature; /* padded */
  double *viscosity;   /* padded (viscous only) */
  /* RANS: cell values accumulated as 1/6 of the six face values
   * (ref: src/procBlock.cpp:1396-1452) */
  double *eddyVisc, *f1, *f2; /* padded; ghosts only across connections */
  double *velGrad;            /* padded, 9: velGrad(r,c) = d u_c / d x_r */
  double *tkeGrad, *omegaGrad; /* ni nj nk 3 */
  double *pressGrad;           /* ni nj nk 3: cell average of the face pressure gradients */
  const double *vol, *fAI, *fAJ, *fAK, *center, *cwI, *cwJ, *cwK, *wallDist;
  int *order; /* hyperplane ordering: 3 ints per cell */
  /* wallData_: one record per face of every viscous-wall surface (NULL for
   * other surfaces); ref: include/wallData.hpp:40-62, src/procBlock.cpp:6287-6290 */
  struct orc_wall_vars **wall;
  /* multigrid (ref: include/gridLevel.hpp:56-59): forcing term of this level (ni nj nk neq,
   * zero on the finest level), update saved before the coarse cycles, and -- kept by the FINE
   * block of a level pair -- coarse cell of every cell (3 ints), its volume weight, the seven
   * trilinear coefficients of its centre in that coarse cell */
  double *forcing, *savedX;
  int *toCoarse;
  double *volFac, *prolong;
} orc_block;

/* wallVars; ref: include/wallData.hpp:40-57 */
typedef struct orc_wall_vars {
  double shearStress[3], heatFlux, yplus, temperature, turbEddyVisc, viscosity,
      density, frictionVelocity, tke, sdr;
} orc_wall_vars;

struct orc_level {
  aither_cfg cfg;
  int neq, ns, nt, asz;
  int nblk;
  orc_block *blk;
  int nconn;
  aither_conn *conn;
};

/* ------------------------------------------------------------------------ */
/* indexing (ref: include/multiArray3d.hpp:96-126)                           */
static inline long cidx(const orc_block *b, int i, int j, int k) {
  return (long)(i + b->g) + (long)(j + b->g) * b->NI +
         (long)(k + b->g) * b->NI * b->NJ;
}
static inline long pidx(const orc_block *b, int i, int j, int k) {
  return (long)i + (long)j * b->ni + (long)k * b->ni * b->nj;
}
/* face arrays have one more entry in their own direction */
static inline long fidxI(const orc_block *b, int i, int j, int k) {
  return (long)(i + b->g) + (long)(j + b->g) * (b->NI + 1) +
         (long)(k + b->g) * (b->NI + 1) * b->NJ;
}
static inline long fidxJ(const orc_block *b, int i, int j, int k) {
  return (long)(i + b->g) + (long)(j + b->g) * b->NI +
         (long)(k + b->g) * b->NI * (b->NJ + 1);
}
static inline long fidxK(const orc_block *b, int i, int j, int k) {
  return (long)(i + b->g) + (long)(j + b->g) * b->NI +
         (long)(k + b->g) * b->NI * b->NJ;
}

/* ------------------------------------------------------------------------ */
/* thermodynamics of a primitive state [rho_s.., u, v, w, p, (k, w)]          */
static double rho_of(const orc_level *h, const double *s) {
  double r = 0.0; /* ref: varArray SpeciesSum, std::accumulate from 0 */
  for (int ss = 0; ss < h->ns; ++ss) r += s[ss];
  return r;
}
/* ref: src/eos.cpp:100-109 */
static double temperature_of(const orc_level *h, const double *s) {
  double rhoR = 0.0;
  for (int ss = 0; ss < h->ns; ++ss) rhoR += s[ss] * h->cfg.gasConstant[ss];
  return s[h->ns + 3] / rhoR;
}
static void mass_fractions(const orc_level *h, const double *s, double *mf) {
  const double rho = rho_of(h, s);
  for (int ss = 0; ss < h->ns; ++ss) mf[ss] = s[ss] / rho;
}
/* ref: src/thermodynamic.cpp:62-82, include/thermodynamic.hpp:55-57,112-119 */
static double cp_mix(const orc_level *h, const double *mf) {
  double cp = 0.0;
  for (int ss = 0; ss < h->ns; ++ss)
    cp += mf[ss] * (h->cfg.gasConstant[ss] * (h->cfg.n[ss] + 1.0));
  return cp;
}
static double cv_mix(const orc_level *h, const double *mf) {
  double cv = 0.0;
  for (int ss = 0; ss < h->ns; ++ss)
    cv += mf[ss] * (h->cfg.gasConstant[ss] * h->cfg.n[ss]);
  return cv;
}
static double gamma_of(const orc_level *h, const double *s) {
  double mf[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  return cp_mix(h, mf) / cv_mix(h, mf);
}
/* ref: include/arrayView.hpp:384-391 */
static double sos_of(const orc_level *h, const double *s) {
  return sqrt(gamma_of(h, s) * s[h->ns + 3] / rho_of(h, s));
}
static double vel_mag(const orc_level *h, const double *s) {
  const double *v = s + h->ns;
  return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}
/* ref: include/arrayView.hpp:400-408, src/eos.cpp:83-88,
 * src/thermodynamic.cpp:95-104 */
static double enthalpy_of(const orc_level *h, const double *s) {
  double mf[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  const double t = temperature_of(h, s);
  double hs = 0.0;
  for (int ss = 0; ss < h->ns; ++ss)
    hs += mf[ss] * (h->cfg.hf[ss] +
                    (h->cfg.gasConstant[ss] * (h->cfg.n[ss] + 1.0)) * t);
  const double vel = vel_mag(h, s);
  return hs + 0.5 * vel * vel;
}
/* ref: include/arrayView.hpp:432-441 */
static double energy_of(const orc_level *h, const double *s) {
  double mf[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  const double t = temperature_of(h, s);
  double e = 0.0;
  for (int ss = 0; ss < h->ns; ++ss)
    e += mf[ss] *
         (h->cfg.hf[ss] + (h->cfg.gasConstant[ss] * h->cfg.n[ss]) * t);
  const double vel = vel_mag(h, s);
  return e + 0.5 * vel * vel;
}
/* ref: include/primitive.hpp:181-199 (PrimToCons) */
static void prim_to_cons(const orc_level *h, const double *s, double *c) {
  const int ns = h->ns;
  for (int ss = 0; ss < ns; ++ss) c[ss] = s[ss];
  const double rho = rho_of(h, s);
  c[ns] = rho * s[ns];
  c[ns + 1] = rho * s[ns + 1];
  c[ns + 2] = rho * s[ns + 2];
  c[ns + 3] = rho * energy_of(h, s);
  for (int tt = 0; tt < h->nt; ++tt) c[ns + 4 + tt] = rho * s[ns + 4 + tt];
}
/* ref: include/primitive.hpp:150-177 (primitive(cons, phys)),
 * src/eos.cpp:40-63, src/thermodynamic.cpp:107-113 */
static void cons_to_prim(const orc_level *h, const double *c, double *s) {
  const int ns = h->ns;
  for (int ss = 0; ss < ns; ++ss) s[ss] = c[ss];
  double rho = 0.0;
  for (int ss = 0; ss < ns; ++ss) rho += c[ss];
  s[ns] = c[ns] / rho;
  s[ns + 1] = c[ns + 1] / rho;
  s[ns + 2] = c[ns + 2] / rho;
  const double energy = c[ns + 3] / rho;
  const double vel = sqrt(s[ns] * s[ns] + s[ns + 1] * s[ns + 1] +
                          s[ns + 2] * s[ns + 2]);
  const double specEnergy = energy - 0.5 * vel * vel;
  double rhoSum = 0.0;
  for (int ss = 0; ss < ns; ++ss) rhoSum += s[ss];
  double mf[AITHER_MAX_SPECIES];
  for (int ss = 0; ss < ns; ++ss) mf[ss] = s[ss] / rhoSum;
  double hf = 0.0;
  for (int ss = 0; ss < ns; ++ss) hf += h->cfg.hf[ss] * mf[ss];
  const double temperature = (specEnergy - hf) / cv_mix(h, mf);
  double p = 0.0;
  for (int ss = 0; ss < ns; ++ss)
    p += s[ss] * h->cfg.gasConstant[ss] * temperature;
  s[ns + 3] = p;
  for (int tt = 0; tt < h->nt; ++tt) {
    const double v = c[ns + 4 + tt] / rho;
    s[ns + 4 + tt] = v > 1.0e-20 ? v : 1.0e-20; /* ref: turbulence.hpp:72-73 */
  }
}
/* ref: include/primitive.hpp:206-231 (UpdatePrimWithCons) */
static void update_prim_with_cons(const orc_level *h, const double *s,
                                  const double *du, double *out) {
  const int ns = h->ns;
  double c[MAXEQ];
  prim_to_cons(h, s, c);
  for (int e = 0; e < h->neq; ++e) c[e] += du[e];
  double rho = 0.0;
  for (int ss = 0; ss < ns; ++ss) rho += c[ss];
  double mf[AITHER_MAX_SPECIES];
  double total = 0.0;
  for (int ss = 0; ss < ns; ++ss) {
    mf[ss] = c[ss] / rho;
    mf[ss] = mf[ss] > 0.0 ? mf[ss] : 0.0;
    total += mf[ss];
  }
  for (int ss = 0; ss < ns; ++ss) mf[ss] /= total;
  for (int ss = 0; ss < ns; ++ss) c[ss] = rho * mf[ss];
  cons_to_prim(h, c, out);
}

/* ------------------------------------------------------------------------ */
/* reconstruction                                                             */
static double limiter_fn(int lim, double r) {
  if (lim == AITHER_LIMITER_VAN_ALBADA) { /* ref: src/limiter.cpp:37-46 */
    const double r2 = r * r;
    const double l = (r + r2) / (1.0 + r2);
    return l > 0.0 ? l : 0.0;
  } else if (lim == AITHER_LIMITER_MINMOD) { /* ref: src/limiter.cpp:24-34 */
    const double m = 1.0 < r ? 1.0 : r;
    return 0.0 > m ? 0.0 : m;
  }
  return 1.0;
}
/* ref: include/reconstruction.hpp:110-154 (FaceReconMUSCL) */
void orc_muscl(const double *uw2, const double *uw1, const double *dw1, int n,
               double kappa, int limiter, double w2, double w1, double wd,
               double *face) {
  const double dPlus = (w1 + w1) / (w1 + wd);
  const double dMinus = (w1 + w1) / (w1 + w2);
  for (int e = 0; e < n; ++e) {
    const double r = (ORC_EPS + (dw1[e] - uw1[e]) * dPlus) /
                     (ORC_EPS + (uw1[e] - uw2[e]) * dMinus);
    double lim = 1.0, invLim = 1.0;
    if (limiter != AITHER_LIMITER_NONE) {
      lim = limiter_fn(limiter, r);
      invLim = limiter_fn(limiter, 1.0 / r);
    }
    face[e] = uw1[e] + 0.25 * ((uw1[e] - uw2[e]) * dMinus) *
                           ((1.0 - kappa) * lim + (1.0 + kappa) * r * invLim);
  }
}

/* ref: include/utility.hpp:103-114 */
static double stencil_width(const double *w, int start, int end) {
  double width = 0.0;
  if (end > start) {
    for (int q = start; q < end; ++q) width += w[q];
  } else if (start > end) {
    for (int q = end; q < start; ++q) width += w[q];
    width = -1.0 * width;
  }
  return width;
}
/* ref: src/utility.cpp:449-483 (LagrangeCoeff) */
static void lagrange_coeff(const double *w, int degree, int rr, int ii,
                           double *coeffs) {
  for (int jj = 0; jj <= degree; ++jj) {
    coeffs[jj] = 0.0;
    for (int mm = jj + 1; mm <= degree + 1; ++mm) {
      double numer = 0.0, denom = 1.0;
      for (int ll = 0; ll <= degree + 1; ++ll) {
        if (ll != mm) {
          double numProd = 1.0;
          for (int qq = 0; qq <= degree + 1; ++qq) {
            if (qq != mm && qq != ll)
              numProd *= stencil_width(w, ii - rr + qq, ii + 1);
          }
          numer += numProd;
          denom *= stencil_width(w, ii - rr + ll, ii - rr + mm);
        }
      }
      coeffs[jj] += numer / denom;
    }
    coeffs[jj] *= w[ii - rr + jj];
  }
}
/* ref: include/utility.hpp:116-122 */
static double deriv2nd(double x0, double x1, double x2, double y0, double y1,
                       double y2) {
  const double fwd = (y2 - y1) / (0.5 * (x2 + x1));
  const double bck = (y1 - y0) / (0.5 * (x1 + x0));
  return (fwd - bck) / (0.25 * (x2 + x0) + 0.5 * x1);
}
/* ref: include/reconstruction.hpp:157-183 */
static double beta_integral1(double d1, double d2, double dx, double x) {
  return ((d1 * d1) * x + d1 * d2 * x * x + (d2 * d2) * pow(x, 3.0) / 3.0) *
             dx +
         (d2 * d2) * x * pow(dx, 3.0);
}
static double beta_integral(double d1, double d2, double dx, double xl,
                            double xh) {
  return beta_integral1(d1, d2, dx, xh) - beta_integral1(d1, d2, dx, xl);
}
/* ref: include/reconstruction.hpp:244-310 (FaceReconWENO); u[0..4] =
 * upwind3, upwind2, upwind1, downwind1, downwind2 */
void orc_weno(const double *u[5], const double w[5], int n, int isWenoZ,
              double *face) {
  double c0[3], c1[3], c2[3], fc[5];
  lagrange_coeff(w, 2, 2, 2, c0);
  lagrange_coeff(w, 2, 1, 2, c1);
  lagrange_coeff(w, 2, 0, 2, c2);
  lagrange_coeff(w, 4, 2, 2, fc);
  const double lw0 = fc[0] / c0[0];
  const double lw1 = fc[4] / c2[2];
  const double lw2 = 1.0 - lw0 - lw1;
  for (int e = 0; e < n; ++e) {
    const double y0 = u[0][e], y1 = u[1][e], y2 = u[2][e], y3 = u[3][e],
                 y4 = u[4][e];
    const double st0 = c0[0] * y0 + c0[1] * y1 + c0[2] * y2;
    const double st1 = c1[0] * y1 + c1[1] * y2 + c1[2] * y3;
    const double st2 = c2[0] * y2 + c2[1] * y3 + c2[2] * y4;
    /* Beta0(uw3, uw2, uw1, ...) ref: reconstruction.hpp:185-202 */
    double d2 = deriv2nd(w[0], w[1], w[2], y0, y1, y2);
    double d1 = (y2 - y1) / (0.5 * (w[2] + w[1])) + 0.5 * w[2] * d2;
    const double b0 = beta_integral(d1, d2, w[2], -0.5 * w[2], 0.5 * w[2]);
    /* Beta1(uw2, uw1, dw1, ...) ref: :204-221 */
    d2 = deriv2nd(w[1], w[2], w[3], y1, y2, y3);
    d1 = (y3 - y2) / (0.5 * (w[3] + w[2])) - 0.5 * w[2] * d2;
    const double b1 = beta_integral(d1, d2, w[2], -0.5 * w[2], 0.5 * w[2]);
    /* Beta2(uw1, dw1, dw2, ...) ref: :223-240 */
    d2 = deriv2nd(w[2], w[3], w[4], y2, y3, y4);
    d1 = (y3 - y2) / (0.5 * (w[3] + w[2])) - 0.5 * w[2] * d2;
    const double b2 = beta_integral(d1, d2, w[2], -0.5 * w[2], 0.5 * w[2]);
    double n0, n1, n2;
    if (isWenoZ) {
      const double tau5 = fabs(b0 - b2);
      const double eps = 1.0e-40;
      double t = tau5 / (eps + b0);
      n0 = lw0 * (1.0 + t * t);
      t = tau5 / (eps + b1);
      n1 = lw1 * (1.0 + t * t);
      t = tau5 / (eps + b2);
      n2 = lw2 * (1.0 + t * t);
    } else {
      const double eps = 1.0e-6;
      n0 = lw0 / ((eps + b0) * (eps + b0));
      n1 = lw1 / ((eps + b1) * (eps + b1));
      n2 = lw2 / ((eps + b2) * (eps + b2));
    }
    const double sum = n0 + n1 + n2;
    n0 /= sum;
    n1 /= sum;
    n2 /= sum;
    face[e] = n0 * st0 + n1 * st1 + n2 * st2;
  }
}

/* ------------------------------------------------------------------------ */
/* inviscid fluxes                                                            */
/* ref: include/inviscidFlux.hpp:128-159 (ConstructFromPrim) */
static void physical_flux(const orc_level *h, const double *s,
                          const double n[3], double *f) {
  const int ns = h->ns;
  const double *v = s + ns;
  const double velNorm = v[0] * n[0] + v[1] * n[1] + v[2] * n[2];
  for (int ss = 0; ss < ns; ++ss) f[ss] = s[ss] * velNorm;
  const double rho = rho_of(h, s);
  const double p = s[ns + 3];
  f[ns] = rho * velNorm * v[0] + p * n[0];
  f[ns + 1] = rho * velNorm * v[1] + p * n[1];
  f[ns + 2] = rho * velNorm * v[2] + p * n[2];
  f[ns + 3] = rho * velNorm * enthalpy_of(h, s);
  for (int tt = 0; tt < h->nt; ++tt)
    f[ns + 4 + tt] = rho * velNorm * s[ns + 4 + tt];
}
/* ref: include/inviscidFlux.hpp:260-382 (RoeFlux),
 * include/primitive.hpp:245-280 (RoeAveragedState) */
static void roe_flux(const orc_level *h, const double *l, const double *r,
                     const double n[3], double *flux) {
  const int ns = h->ns, nt = h->nt, neq = h->neq;
  const int imx = ns, imy = ns + 1, imz = ns + 2, ie = ns + 3, it = ns + 4;
  double roe[MAXEQ];
  const double denRatio = sqrt(rho_of(h, r) / rho_of(h, l));
  for (int ss = 0; ss < ns; ++ss) roe[ss] = l[ss] * denRatio;
  roe[imx] = (l[imx] + denRatio * r[imx]) / (1.0 + denRatio);
  roe[imy] = (l[imy] + denRatio * r[imy]) / (1.0 + denRatio);
  roe[imz] = (l[imz] + denRatio * r[imz]) / (1.0 + denRatio);
  roe[ie] = (l[ie] + denRatio * r[ie]) / (1.0 + denRatio);
  for (int tt = 0; tt < nt; ++tt)
    roe[it + tt] = (l[it + tt] + denRatio * r[it + tt]) / (1.0 + denRatio);

  const double hR = enthalpy_of(h, roe);
  const double aR = sos_of(h, roe);
  const double rhoR = rho_of(h, roe);
  const double velNormR = roe[imx] * n[0] + roe[imy] * n[1] + roe[imz] * n[2];
  double mfR[AITHER_MAX_SPECIES];
  mass_fractions(h, roe, mfR);
  double delta[MAXEQ];
  for (int e = 0; e < neq; ++e) delta[e] = r[e] - l[e];
  double deltaRho = 0.0;
  for (int ss = 0; ss < ns; ++ss) deltaRho += delta[ss];
  const double normVelDiff =
      delta[imx] * n[0] + delta[imy] * n[1] + delta[imz] * n[2];

  double diss[MAXEQ];
  for (int e = 0; e < neq; ++e) diss[e] = 0.0;

  /* left moving acoustic wave */
  double waveSpeed = fabs(velNormR - aR);
  const double entropyFix = 0.1;
  if (waveSpeed < entropyFix)
    waveSpeed = 0.5 * (waveSpeed * waveSpeed / entropyFix + entropyFix);
  double waveStrength = (delta[ie] - rhoR * aR * normVelDiff) / (2.0 * aR * aR);
  double wss = waveSpeed * waveStrength;
  for (int ss = 0; ss < ns; ++ss) diss[ss] += wss * mfR[ss];
  diss[imx] += wss * (roe[imx] - aR * n[0]);
  diss[imy] += wss * (roe[imy] - aR * n[1]);
  diss[imz] += wss * (roe[imz] - aR * n[2]);
  diss[ie] += wss * (hR - aR * velNormR);
  for (int tt = 0; tt < nt; ++tt) diss[it + tt] += wss * roe[it + tt];

  /* entropy wave */
  waveSpeed = fabs(velNormR);
  for (int ss = 0; ss < ns; ++ss) {
    waveStrength = -delta[ie] / (aR * aR);
    wss = waveSpeed * waveStrength;
    diss[ss] += wss * mfR[ss] + waveSpeed * delta[ss];
  }
  waveStrength = deltaRho - delta[ie] / (aR * aR);
  wss = waveSpeed * waveStrength;
  diss[imx] += wss * roe[imx];
  diss[imy] += wss * roe[imy];
  diss[imz] += wss * roe[imz];
  diss[ie] += wss * 0.5 *
              (roe[imx] * roe[imx] + roe[imy] * roe[imy] + roe[imz] * roe[imz]);

  /* shear wave */
  waveStrength = rhoR;
  wss = waveSpeed * waveStrength;
  diss[imx] += wss * (delta[imx] - normVelDiff * n[0]);
  diss[imy] += wss * (delta[imy] - normVelDiff * n[1]);
  diss[imz] += wss * (delta[imz] - normVelDiff * n[2]);
  diss[ie] += wss * ((roe[imx] * delta[imx] + roe[imy] * delta[imy] +
                      roe[imz] * delta[imz]) -
                     velNormR * normVelDiff);

  /* right moving acoustic wave */
  waveSpeed = fabs(velNormR + aR);
  if (waveSpeed < entropyFix)
    waveSpeed = 0.5 * (waveSpeed * waveSpeed / entropyFix + entropyFix);
  waveStrength = (delta[ie] + rhoR * aR * normVelDiff) / (2.0 * aR * aR);
  wss = waveSpeed * waveStrength;
  for (int ss = 0; ss < ns; ++ss) diss[ss] += wss * mfR[ss];
  diss[imx] += wss * (roe[imx] + aR * n[0]);
  diss[imy] += wss * (roe[imy] + aR * n[1]);
  diss[imz] += wss * (roe[imz] + aR * n[2]);
  diss[ie] += wss * (hR + aR * velNormR);
  for (int tt = 0; tt < nt; ++tt) diss[it + tt] += wss * roe[it + tt];

  /* turbulence waves */
  if (nt > 0) {
    waveSpeed = fabs(velNormR);
    for (int tt = 0; tt < nt; ++tt) {
      waveStrength = rhoR * delta[it + tt] + roe[it + tt] * deltaRho -
                     delta[ie] * roe[it + tt] / (aR * aR);
      wss = waveSpeed * waveStrength;
      diss[it + tt] += wss * 1.0;
    }
  }

  double fl[MAXEQ], fr[MAXEQ];
  physical_flux(h, l, n, fl);
  physical_flux(h, r, n, fr);
  /* ref: src/inviscidFlux.cpp:27-33: left += right - diss; left *= 0.5 */
  for (int e = 0; e < neq; ++e) flux[e] = (fl[e] + (fr[e] - diss[e])) * 0.5;
}

static double sign_of(double v) { return (double)((0.0 < v) - (v < 0.0)); }

/* ref: include/inviscidFlux.hpp:396-481 (AUSMFlux), :161-208 */
static void ausm_flux(const orc_level *h, const double *l, const double *r,
                      const double n[3], double *f) {
  const int ns = h->ns, nt = h->nt;
  const int imx = ns, imy = ns + 1, imz = ns + 2, ie = ns + 3, it = ns + 4;
  const double velNormL = l[imx] * n[0] + l[imy] * n[1] + l[imz] * n[2];
  const double velNormR = r[imx] * n[0] + r[imy] * n[1] + r[imz] * n[2];
  const double sosL = sos_of(h, l);
  const double sosR = sos_of(h, r);
  const double sosStar = sqrt(sosL * sosR);
  const double vel = 0.5 * (velNormL + velNormR);
  double sos = sosStar;
  if (vel < 0.0) {
    sos = sosStar * sosStar / (velNormR > sosStar ? velNormR : sosStar);
  } else if (vel > 0.0) {
    sos = sosStar * sosStar / (velNormL > sosStar ? velNormL : sosStar);
  }
  const double ml = velNormL / sos;
  const double mr = velNormR / sos;
  const double mPlusL = fabs(ml) <= 1.0 ? 0.25 * pow(ml + 1.0, 2.0)
                                        : 0.5 * (ml + fabs(ml));
  const double mMinusR = fabs(mr) <= 1.0 ? -0.25 * pow(mr - 1.0, 2.0)
                                         : 0.5 * (mr - fabs(mr));
  const double pPlus = fabs(ml) <= 1.0
                           ? 0.25 * pow(ml + 1.0, 2.0) * (2.0 - ml)
                           : 0.5 * (1.0 + sign_of(ml));
  const double pMinus = fabs(mr) <= 1.0
                            ? 0.25 * pow(mr - 1.0, 2.0) * (2.0 + mr)
                            : 0.5 * (1.0 - sign_of(mr));
  const double pl = l[ie], pr = r[ie];
  const double ps = pPlus * pl + pMinus * pr;
  const double ratio = pl / pr < pr / pl ? pl / pr : pr / pl;
  const double w = 1.0 - pow(ratio, 3.0);
  const double fl = fabs(ml) < 1.0 ? pl / ps - 1.0 : 0.0;
  const double fr = fabs(mr) < 1.0 ? pr / ps - 1.0 : 0.0;
  const double mavg = mPlusL + mMinusR;
  const double mPlusLBar =
      mavg >= 0.0 ? mPlusL + mMinusR * ((1.0 - w) * (1.0 + fr) - fl)
                  : mPlusL * w * (1.0 + fl);
  const double mMinusRBar =
      mavg >= 0.0 ? mMinusR * w * (1.0 + fr)
                  : mMinusR + mPlusL * ((1.0 - w) * (1.0 + fl) - fr);

  const double vl = mPlusLBar * sos;
  for (int ss = 0; ss < ns; ++ss) f[ss] = l[ss] * vl;
  const double rhoL = rho_of(h, l);
  f[imx] = rhoL * vl * l[imx] + pPlus * pl * n[0];
  f[imy] = rhoL * vl * l[imy] + pPlus * pl * n[1];
  f[imz] = rhoL * vl * l[imz] + pPlus * pl * n[2];
  f[ie] = rhoL * vl * enthalpy_of(h, l);
  for (int tt = 0; tt < nt; ++tt) f[it + tt] = rhoL * vl * l[it + tt];
  const double vr = mMinusRBar * sos;
  for (int ss = 0; ss < ns; ++ss) f[ss] += r[ss] * vr;
  const double rhoR = rho_of(h, r);
  f[imx] += rhoR * vr * r[imx] + pMinus * pr * n[0];
  f[imy] += rhoR * vr * r[imy] + pMinus * pr * n[1];
  f[imz] += rhoR * vr * r[imz] + pMinus * pr * n[2];
  f[ie] += rhoR * vr * enthalpy_of(h, r);
  for (int tt = 0; tt < nt; ++tt) f[it + tt] += rhoR * vr * r[it + tt];
}

static void inviscid_flux(const orc_level *h, const double *l, const double *r,
                          const double n[3], double *f) {
  if (h->cfg.invFlux == AITHER_FLUX_ROE) roe_flux(h, l, r, n, f);
  else ausm_flux(h, l, r, n, f);
}

/* ref: include/spectralRadius.hpp:44-65 (InvCellSpectralRadius) */
static double inv_cell_spec_rad(const orc_level *h, const double *s,
                                const double *fL, const double *fR) {
  double a[3] = {0.5 * (fL[0] + fR[0]), 0.5 * (fL[1] + fR[1]),
                 0.5 * (fL[2] + fR[2])};
  const double mag = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  a[0] /= mag;
  a[1] /= mag;
  a[2] /= mag;
  const double fMag = 0.5 * (fL[3] + fR[3]);
  const double *v = s + h->ns;
  return (fabs(v[0] * a[0] + v[1] * a[1] + v[2] * a[2]) + sos_of(h, s)) * fMag;
}
/* ref: include/spectralRadius.hpp:67-80 (InvFaceSpectralRadius) */
static double inv_face_spec_rad(const orc_level *h, const double *s,
                                const double *fA) {
  const double *v = s + h->ns;
  return 0.5 * fA[3] *
         (fabs(v[0] * fA[0] + v[1] * fA[1] + v[2] * fA[2]) + sos_of(h, s));
}

/* ------------------------------------------------------------------------ */
/* boundary conditions                                                        */
static const aither_bc_state *bc_data(const orc_level *h, int tag) {
  for (int b = 0; b < h->cfg.numBCStates; ++b)
    if (h->cfg.bcStates[b].tag == tag) return &h->cfg.bcStates[b];
  fprintf(stderr, "oracle: no BC state with tag %d\n", tag);
  abort();
  return NULL;
}
/* ref: src/ghostStates.cpp:691-708 (ExtrapolateHoldMixture) */
static void extrapolate_hold_mixture(const orc_level *h, const double *bnd,
                                     double factor, const double *interior,
                                     double *out) {
  const int ns = h->ns;
  const double bndRho = rho_of(h, bnd);
  double bndMf[AITHER_MAX_SPECIES];
  mass_fractions(h, bnd, bndMf);
  const double intRho = rho_of(h, interior);
  const double ghostRho = factor * bndRho - intRho;
  if (ghostRho <= 0.0) {
    for (int e = 0; e < h->neq; ++e) out[e] = bnd[e];
    return;
  }
  double tmp[MAXEQ];
  for (int e = 0; e < h->neq; ++e) tmp[e] = factor * bnd[e] - interior[e];
  for (int ss = 0; ss < ns; ++ss) {
    const double v = ghostRho * bndMf[ss];
    tmp[ss] = v > 0.0 ? v : 0.0;
  }
  for (int e = 0; e < h->neq; ++e) out[e] = tmp[e];
}

/* ref: src/ghostStates.cpp:62-689 (GetGhostState), inviscid / low-Re subset */
static double eff_conductivity(const orc_level *h, double t, const double *s);
static double viscosity_of(const orc_level *h, double t, const double *s);
/* primitive::ApplyFarfieldTurbBC; ref: src/primitive.cpp:83-98 */
static void apply_farfield_turb(const orc_level *h, double *s, const double vel[3],
                                double turbInten, double viscRatio) {
  const int it = h->ns + 4;
  const double vmag = sqrt(vel[0] * vel[0] + vel[1] * vel[1] + vel[2] * vel[2]);
  s[it] = 1.5 * pow(turbInten * vmag, 2.0);
  s[it + 1] = rho_of(h, s) * s[it] / (viscRatio * viscosity_of(h, temperature_of(h, s), s));
  for (int tt = 0; tt < h->nt; ++tt)
    s[it + tt] = s[it + tt] > 1.0e-20 ? s[it + tt] : 1.0e-20;
}
/* what the implicit time integrators hand to GetGhostState besides the interior state
 * (ref: src/procBlock.cpp:2506-2522, :6272-6285): time step, state at time n and gradients of
 * the boundary-adjacent cell, average / maximum normal Mach number of the surface patch. Set by
 * assign_inviscid_ghosts around its ghost_state calls; read by the non-reflecting branches. */
typedef struct {
  double dt, stateN[16], pressGrad[3], velGrad[9], avgMach, maxMach;
} orc_bc_extra;
static const orc_bc_extra *g_bc_extra = NULL;
static void wall_law_eval(const orc_level *h, const aither_bc_state *bc, int mode,
                          const double *state, double wallDist, const double area[3],
                          int isLower, orc_wall_vars *wv);
static double cp_mix(const orc_level *h, const double *mf);
static double turb_prandtl(const orc_level *h);
static void ghost_state(const orc_level *h, const double *interior, int bcType,
                        const double areaVec[3], int surf, int tag, int layer,
                        double wallDist, double nuW, double *ghost, orc_wall_vars *wvOut) {
  const int rans = h->nt > 0, it = h->ns + 4;
  const int ns = h->ns, neq = h->neq;
  const int imx = ns, imy = ns + 1, imz = ns + 2, ie = ns + 3;
  for (int e = 0; e < neq; ++e) ghost[e] = interior[e];
  const int isLower = surf % 2 == 1;
  double nA[3];
  for (int d = 0; d < 3; ++d) nA[d] = isLower ? -1.0 * areaVec[d] : areaVec[d];

  if (bcType == AITHER_BC_SLIP_WALL) { /* ref: :118-133 */
    const double vn =
        interior[imx] * nA[0] + interior[imy] * nA[1] + interior[imz] * nA[2];
    ghost[imx] = interior[imx] - 2.0 * nA[0] * vn;
    ghost[imy] = interior[imy] - 2.0 * nA[1] * vn;
    ghost[imz] = interior[imz] - 2.0 * nA[2] * vn;
  } else if (bcType == AITHER_BC_VISCOUS_WALL) { /* ref: :134-258, low-Re */
    const aither_bc_state *bc = bc_data(h, tag);
    const double zero3[3] = {0.0, 0.0, 0.0};
    const double *velWall = bc ? bc->velocity : zero3;
    ghost[imx] = 2.0 * velWall[0] - interior[imx];
    ghost[imy] = 2.0 * velWall[1] - interior[imy];
    ghost[imz] = 2.0 * velWall[2] - interior[imz];
    orc_wall_vars wv;
    memset(&wv, 0, sizeof(wv));
    const int wallLaw = bc && bc->isWallLaw;
    int lowRe = 1; /* low-Re treatment, or the wall law handed over to it (y+ < 10) */
    if (bc && (bc->isIsothermal || bc->isConstantHeatFlux)) {
      double mf[AITHER_MAX_SPECIES];
      mass_fractions(h, interior, mf);
      double tGhost;
      if (wallLaw) { /* ref: :149-181 (isothermal), :195-224 (heat flux) */
        wall_law_eval(h, bc, bc->isIsothermal ? 2 : 1, interior, wallDist, nA, isLower, &wv);
        lowRe = wv.yplus < 10.0; /* wallVars::SwitchToLowRe, include/wallData.hpp:57 */
      }
      if (bc->isIsothermal) {
        if (!lowRe) { /* ref: :160-171 */
          const double kappa =
              eff_conductivity(h, wv.temperature, interior) +
              wv.turbEddyVisc * cp_mix(h, mf) / turb_prandtl(h);
          tGhost = bc->temperature - wv.heatFlux / kappa * 2.0 * wallDist;
        } else { /* ref: :183-189 */
          tGhost = 2.0 * bc->temperature - temperature_of(h, interior);
        }
      } else {
        if (!lowRe) { /* ref: :212-219 */
          tGhost = 2.0 * wv.temperature - temperature_of(h, interior);
        } else { /* ref: :226-236 */
          const double t = temperature_of(h, interior);
          const double kappa = eff_conductivity(h, t, interior);
          tGhost = temperature_of(h, interior) -
                   bc->heatFlux / kappa * 2.0 * wallDist;
        }
      }
      double R = 0.0; /* ref: src/eos.cpp:111-115 (DensityTP) */
      for (int ss = 0; ss < ns; ++ss) R += mf[ss] * h->cfg.gasConstant[ss];
      const double rho = ghost[ie] / (R * tGhost);
      for (int ss = 0; ss < ns; ++ss) ghost[ss] = rho * mf[ss];
    } else if (wallLaw) { /* adiabatic; ref: :243-256 */
      wall_law_eval(h, bc, 0, interior, wallDist, nA, isLower, &wv);
      lowRe = wv.yplus < 10.0;
    }
    if (rans && !lowRe) { /* wall-law k and omega; ref: :173-180, :221-228, :248-255 */
      ghost[it] = 2.0 * wv.tke - interior[it];
      ghost[it + 1] = 2.0 * wv.sdr - interior[it + 1];
      if (layer > 1) {
        ghost[it] = layer * ghost[it] - wv.tke;
        ghost[it + 1] = layer * ghost[it + 1] - wv.sdr;
      }
    }
    if (wvOut) *wvOut = wv;
    if (rans && lowRe) { /* ref: :262-281 (low-Re wall) */
      ghost[it] = -1.0 * interior[it];
      const double scaling = h->cfg.nondimScaling;
      const double wWall = scaling * scaling * 60.0 * nuW /
                           (wallDist * wallDist * (h->cfg.turbModel == AITHER_TURB_SST
                                                       ? 0.075 : 0.0708));
      ghost[it + 1] = 2.0 * wWall - interior[it + 1];
      if (layer > 1) ghost[it + 1] = layer * ghost[it + 1] - wWall;
    }
  } else if (bcType == AITHER_BC_CHARACTERISTIC) { /* ref: :289-386 */
    const aither_bc_state *bc = bc_data(h, tag);
    double freeState[MAXEQ];
    for (int e = 0; e < neq; ++e) freeState[e] = 0.0;
    for (int ss = 0; ss < ns; ++ss)
      freeState[ss] = bc->density * bc->massFractions[ss];
    freeState[imx] = bc->velocity[0];
    freeState[imy] = bc->velocity[1];
    freeState[imz] = bc->velocity[2];
    freeState[ie] = bc->pressure;
    const double velIntNorm =
        interior[imx] * nA[0] + interior[imy] * nA[1] + interior[imz] * nA[2];
    const double SoSInt = sos_of(h, interior);
    const double machInt = fabs(velIntNorm) / SoSInt;
    if (machInt >= 1.0 && velIntNorm < 0.0) {
      for (int e = 0; e < neq; ++e) ghost[e] = freeState[e];
      if (rans) apply_farfield_turb(h, ghost, bc->velocity, bc->turbulenceIntensity,
                                    bc->eddyViscosityRatio);
    } else if (machInt >= 1.0 && velIntNorm >= 0.0) {
      /* supersonic outflow: interior */
    } else if (machInt < 1.0 && velIntNorm < 0.0) {
      const double rhoSoSInt = rho_of(h, interior) * SoSInt;
      const double vd[3] = {freeState[imx] - interior[imx],
                            freeState[imy] - interior[imy],
                            freeState[imz] - interior[imz]};
      ghost[ie] = 0.5 * (freeState[ie] + interior[ie] -
                         rhoSoSInt * (nA[0] * vd[0] + nA[1] * vd[1] +
                                      nA[2] * vd[2]));
      const double deltaPressure = freeState[ie] - ghost[ie];
      const double rho = rho_of(h, freeState) - deltaPressure / (SoSInt * SoSInt);
      double fmf[AITHER_MAX_SPECIES];
      mass_fractions(h, freeState, fmf);
      for (int ss = 0; ss < ns; ++ss) ghost[ss] = rho * fmf[ss];
      ghost[imx] = freeState[imx] - nA[0] * deltaPressure / rhoSoSInt;
      ghost[imy] = freeState[imy] - nA[1] * deltaPressure / rhoSoSInt;
      ghost[imz] = freeState[imz] - nA[2] * deltaPressure / rhoSoSInt;
      if (rans) apply_farfield_turb(h, ghost, bc->velocity, bc->turbulenceIntensity,
                                    bc->eddyViscosityRatio);
    } else if (machInt < 1.0 && velIntNorm >= 0.0) {
      const double rhoSoSInt = rho_of(h, interior) * SoSInt;
      const double deltaPressure = interior[ie] - freeState[ie];
      const double rho = rho_of(h, interior) - deltaPressure / (SoSInt * SoSInt);
      double imf[AITHER_MAX_SPECIES];
      mass_fractions(h, interior, imf);
      for (int ss = 0; ss < ns; ++ss) ghost[ss] = rho * imf[ss];
      ghost[imx] = interior[imx] + nA[0] * deltaPressure / rhoSoSInt;
      ghost[imy] = interior[imy] + nA[1] * deltaPressure / rhoSoSInt;
      ghost[imz] = interior[imz] + nA[2] * deltaPressure / rhoSoSInt;
      ghost[ie] = freeState[ie];
    }
    double tmp[MAXEQ];
    extrapolate_hold_mixture(h, ghost, 2.0, interior, tmp);
    for (int e = 0; e < neq; ++e) ghost[e] = tmp[e];
    if (layer > 1) {
      extrapolate_hold_mixture(h, ghost, (double)layer, interior, tmp);
      for (int e = 0; e < neq; ++e) ghost[e] = tmp[e];
      if (rans) apply_farfield_turb(h, ghost, bc->velocity, bc->turbulenceIntensity,
                                    bc->eddyViscosityRatio);
    }
  } else if (bcType == AITHER_BC_INLET) { /* ref: :391-484 */
    const aither_bc_state *bc = bc_data(h, tag);
    double freeState[MAXEQ];
    for (int e = 0; e < neq; ++e) freeState[e] = 0.0;
    for (int ss = 0; ss < ns; ++ss)
      freeState[ss] = bc->density * bc->massFractions[ss];
    freeState[imx] = bc->velocity[0];
    freeState[imy] = bc->velocity[1];
    freeState[imz] = bc->velocity[2];
    freeState[ie] = bc->pressure;
    const double velIntNorm =
        interior[imx] * nA[0] + interior[imy] * nA[1] + interior[imz] * nA[2];
    const double SoSInt = sos_of(h, interior);
    const double machInt = fabs(velIntNorm) / SoSInt;
    if (machInt >= 1.0) {
      for (int e = 0; e < neq; ++e) ghost[e] = freeState[e];
      if (rans) apply_farfield_turb(h, ghost, bc->velocity, bc->turbulenceIntensity,
                                    bc->eddyViscosityRatio);
    } else {
      const double rhoSoSInt = rho_of(h, interior) * SoSInt;
      const double vd[3] = {freeState[imx] - interior[imx],
                            freeState[imy] - interior[imy],
                            freeState[imz] - interior[imz]};
      ghost[ie] = 0.5 * (freeState[ie] + interior[ie] -
                         rhoSoSInt * (nA[0] * vd[0] + nA[1] * vd[1] +
                                      nA[2] * vd[2]));
      double fmf[AITHER_MAX_SPECIES];
      mass_fractions(h, freeState, fmf);
      if (bc->isNonreflecting && g_bc_extra) { /* LODI relaxation; ref: :435-466 */
        const orc_bc_extra *ex = g_bc_extra;
        const double *stateN = ex->stateN;
        const double sigma = 0.25, dt = ex->dt;
        const double rhoN = rho_of(h, stateN);
        const double sosN = sos_of(h, stateN);
        const double rhoSoSN = rhoN * sosN;
        const double deltaPressure = ghost[ie] - stateN[ie];
        const double length = bc->lengthScale;
        const double alpha = sigma * sosN / length;
        const double rhoNp1 = (rhoN + dt * alpha * rho_of(h, freeState) +
                               deltaPressure / (sosN * sosN)) /
                              (1.0 + dt * alpha);
        for (int ss = 0; ss < ns; ++ss) ghost[ss] = rhoNp1 * fmf[ss];
        const double k = alpha * (1.0 - ex->maxMach * ex->maxMach);
        for (int dd = 0; dd < 3; ++dd)
          ghost[imx + dd] = (stateN[imx + dd] + dt * k * freeState[imx + dd] -
                             nA[dd] * deltaPressure / rhoSoSN) /
                            (1.0 + dt * k);
      } else {
      const double deltaPressure = freeState[ie] - ghost[ie];
      const double rho = rho_of(h, freeState) - deltaPressure / (SoSInt * SoSInt);
      for (int ss = 0; ss < ns; ++ss) ghost[ss] = rho * fmf[ss];
      ghost[imx] = freeState[imx] - nA[0] * deltaPressure / rhoSoSInt;
      ghost[imy] = freeState[imy] - nA[1] * deltaPressure / rhoSoSInt;
      ghost[imz] = freeState[imz] - nA[2] * deltaPressure / rhoSoSInt;
      }
      if (rans) apply_farfield_turb(h, ghost, bc->velocity, bc->turbulenceIntensity,
                                    bc->eddyViscosityRatio);
      double tmp[MAXEQ];
      extrapolate_hold_mixture(h, ghost, 2.0, interior, tmp);
      for (int e = 0; e < neq; ++e) ghost[e] = tmp[e];
      if (layer > 1) {
        extrapolate_hold_mixture(h, ghost, (double)layer, interior, tmp);
        for (int e = 0; e < neq; ++e) ghost[e] = tmp[e];
      }
    }
  } else if (bcType == AITHER_BC_SUPERSONIC_INFLOW) { /* ref: :490-515 */
    const aither_bc_state *bc = bc_data(h, tag);
    for (int ss = 0; ss < ns; ++ss)
      ghost[ss] = bc->density * bc->massFractions[ss];
    ghost[imx] = bc->velocity[0];
    ghost[imy] = bc->velocity[1];
    ghost[imz] = bc->velocity[2];
    ghost[ie] = bc->pressure;
    if (rans) apply_farfield_turb(h, ghost, bc->velocity, bc->turbulenceIntensity,
                                  bc->eddyViscosityRatio);
  } else if (bcType == AITHER_BC_SUPERSONIC_OUTFLOW) { /* ref: :522-527 */
    if (layer > 1)
      for (int e = 0; e < neq; ++e) ghost[e] = layer * ghost[e] - interior[e];
  } else if (bcType == AITHER_BC_STAGNATION_INLET) { /* ref: :533-598 */
    const aither_bc_state *bc = bc_data(h, tag);
    const double gam = gamma_of(h, interior);
    const double g = gam - 1.0;
    const double vn =
        interior[imx] * nA[0] + interior[imy] * nA[1] + interior[imz] * nA[2];
    const double sosI = sos_of(h, interior);
    const double rNeg = vn - 2.0 * sosI / g;
    const double vmag = vel_mag(h, interior);
    const double cosTheta = -1.0 * vn / vmag;
    const double magSq = interior[imx] * interior[imx] +
                         interior[imy] * interior[imy] +
                         interior[imz] * interior[imz];
    const double stagSoSsq = pow(sosI, 2.0) + 0.5 * g * magSq;
    const double sosB =
        -1.0 * rNeg * g / (g * cosTheta * cosTheta + 2.0) *
        (1.0 + cosTheta * sqrt((g * cosTheta * cosTheta + 2.0) * stagSoSsq /
                                   (g * rNeg * rNeg) -
                               0.5 * g));
    const double tb = bc->stagnationTemperature * (sosB * sosB / stagSoSsq);
    const double pb =
        bc->stagnationPressure * pow(sosB * sosB / stagSoSsq, gam / g);
    const double vbMag = sqrt(2.0 / g * (bc->stagnationTemperature - tb));
    /* ref: src/eos.cpp:111-115 DensityTP with MixtureGasConstant (inner
     * product from 0) */
    double Rmix = 0.0;
    for (int ss = 0; ss < ns; ++ss)
      Rmix += bc->massFractions[ss] * h->cfg.gasConstant[ss];
    const double rhoGhost = pb / (Rmix * tb);
    for (int ss = 0; ss < ns; ++ss) ghost[ss] = rhoGhost * bc->massFractions[ss];
    ghost[imx] = vbMag * bc->direction[0];
    ghost[imy] = vbMag * bc->direction[1];
    ghost[imz] = vbMag * bc->direction[2];
    ghost[ie] = pb;
    if (rans) apply_farfield_turb(h, ghost, ghost + imx, bc->turbulenceIntensity,
                                  bc->eddyViscosityRatio);
    double tmp[MAXEQ];
    extrapolate_hold_mixture(h, ghost, 2.0, interior, tmp);
    for (int e = 0; e < neq; ++e) ghost[e] = tmp[e];
    if (layer > 1) {
      extrapolate_hold_mixture(h, ghost, (double)layer, interior, tmp);
      for (int e = 0; e < neq; ++e) ghost[e] = tmp[e];
      if (rans) {
        const double gv[3] = {ghost[imx], ghost[imy], ghost[imz]};
        apply_farfield_turb(h, ghost, gv, bc->turbulenceIntensity, bc->eddyViscosityRatio);
      }
    }
  } else if (bcType == AITHER_BC_PRESSURE_OUTLET) { /* ref: :604-664 */
    const aither_bc_state *bc = bc_data(h, tag);
    const double pb = bc->pressure;
    const double SoSInt = sos_of(h, interior);
    const double rhoSoSInt = rho_of(h, interior) * SoSInt;
    ghost[ie] = pb;
    if (bc->isNonreflecting && g_bc_extra) { /* LODI with transverse terms; ref: :614-643 */
      const orc_bc_extra *ex = g_bc_extra;
      const double *stateN = ex->stateN;
      const double dt = ex->dt;
      const double deltaVel = (interior[imx] - stateN[imx]) * nA[0] +
                              (interior[imy] - stateN[imy]) * nA[1] +
                              (interior[imz] - stateN[imz]) * nA[2];
      const double sigma = 0.25;
      const double rhoN = rho_of(h, stateN);
      const double sosN = sos_of(h, stateN);
      const double rhoSoSN = rhoN * sosN;
      const double length = bc->lengthScale;
      const double k = sigma * sosN * (1.0 - ex->maxMach * ex->maxMach) / length;
      const double beta = ex->avgMach;
      const double pgn = ex->pressGrad[0] * nA[0] + ex->pressGrad[1] * nA[1] +
                         ex->pressGrad[2] * nA[2];
      const double vnN = stateN[imx] * nA[0] + stateN[imy] * nA[1] + stateN[imz] * nA[2];
      double pGradT[3], velT[3], vgT[9], dVelN[3];
      for (int dd = 0; dd < 3; ++dd) {
        pGradT[dd] = ex->pressGrad[dd] - pgn * nA[dd];
        velT[dd] = stateN[imx + dd] - vnN * nA[dd];
      }
      /* tensor::RemoveComponent: every row loses its component along the normal
       * (include/tensor.hpp:371-379) */
      for (int r = 0; r < 3; ++r) {
        const double *row = ex->velGrad + 3 * r;
        const double rn = row[0] * nA[0] + row[1] * nA[1] + row[2] * nA[2];
        for (int c = 0; c < 3; ++c) vgT[3 * r + c] = row[c] - rn * nA[c];
      }
      /* tensor::LinearCombination: rows scaled by the normal and summed (:384-389) */
      for (int c = 0; c < 3; ++c) {
        double comb = vgT[c] * nA[0];
        comb += vgT[3 + c] * nA[1];
        comb += vgT[6 + c] * nA[2];
        dVelN[c] = comb;
      }
      double vgSum = 0.0;
      for (int q = 0; q < 9; ++q) vgSum += vgT[q];
      double dnSum = 0.0;
      for (int c = 0; c < 3; ++c) dnSum += dVelN[c];
      const double dVelT_dTrans = vgSum - dnSum;
      const double gam = gamma_of(h, stateN);
      const double dotT = velT[0] * (pGradT[0] - rhoSoSN * dVelN[0]) +
                          velT[1] * (pGradT[1] - rhoSoSN * dVelN[1]) +
                          velT[2] * (pGradT[2] - rhoSoSN * dVelN[2]);
      const double trans = -0.5 * (dotT + gam * stateN[ie] * dVelT_dTrans);
      ghost[ie] = (stateN[ie] + rhoSoSN * deltaVel + dt * k * pb - dt * beta * trans) /
                  (1.0 + dt * k);
    }
    const double deltaPressure = interior[ie] - ghost[ie];
    const double rho = rho_of(h, interior) - deltaPressure / (SoSInt * SoSInt);
    double imf[AITHER_MAX_SPECIES];
    mass_fractions(h, interior, imf);
    for (int ss = 0; ss < ns; ++ss) ghost[ss] = rho * imf[ss];
    ghost[imx] = interior[imx] + nA[0] * deltaPressure / rhoSoSInt;
    ghost[imy] = interior[imy] + nA[1] * deltaPressure / rhoSoSInt;
    ghost[imz] = interior[imz] + nA[2] * deltaPressure / rhoSoSInt;
    const double gvn = ghost[imx] * nA[0] + ghost[imy] * nA[1] + ghost[imz] * nA[2];
    if (gvn / sos_of(h, ghost) >= 1.0)
      for (int e = 0; e < neq; ++e) ghost[e] = interior[e];
    for (int e = 0; e < neq; ++e) ghost[e] = 2.0 * ghost[e] - interior[e];
    if (layer > 1)
      for (int e = 0; e < neq; ++e) ghost[e] = layer * ghost[e] - interior[e];
  } else if (bcType == AITHER_BC_INTERBLOCK || bcType == AITHER_BC_PERIODIC) {
    /* filled by the halo swap */
  } else {
    fprintf(stderr, "oracle: BC type %d not supported\n", bcType);
    abort();
  }
}

static void surface_dirs(int surfType, int *d3) {
  *d3 = (surfType - 1) / 2; /* 0 = i, 1 = j, 2 = k */
}
static int surface_type(const aither_surface *s) {
  /* ref: src/boundaryConditions.cpp:2424-2452 */
  if (s->imin == s->imax) return s->imax == 0 ? 1 : 2;
  if (s->jmin == s->jmax) return s->jmax == 0 ? 3 : 4;
  return s->kmax == 0 ? 5 : 6;
}

/* ref: src/procBlock.cpp:2449-2530 (AssignInviscidGhostCells) */
static void assign_inviscid_ghosts(orc_level *h, orc_block *b) {
  const int neq = h->neq;
  for (int layer = 1; layer <= b->g; ++layer) {
    for (int s = 0; s < b->nsurf; ++s) {
      const aither_surface *sf = &b->surf[s];
      if (sf->type == AITHER_BC_INTERBLOCK || sf->type == AITHER_BC_PERIODIC)
        continue;
      const int st = surface_type(sf);
      int d3;
      surface_dirs(st, &d3);
      const int nd[3] = {b->ni, b->nj, b->nk};
      const int r3 = d3 == 0 ? sf->imin : (d3 == 1 ? sf->jmin : sf->kmin);
      int gCell, iCell, aCell;
      const int bnd = r3;
      if (st % 2 == 0) {
        gCell = r3 + layer - 1;
        iCell = r3 - layer;
        aCell = r3 - 1;
        if (iCell < 0) iCell = 0;
      } else {
        gCell = r3 - layer;
        iCell = r3 + layer - 1;
        aCell = r3;
        if (iCell >= nd[d3]) iCell = nd[d3] - 1;
      }
      int bcType = sf->type;
      if (bcType == AITHER_BC_VISCOUS_WALL) bcType = AITHER_BC_SLIP_WALL;
      const int src = bcType == AITHER_BC_SLIP_WALL ? iCell : aCell;
      /* tangential cell ranges */
      int lo[3] = {sf->imin, sf->jmin, sf->kmin};
      int hi[3] = {sf->imax, sf->jmax, sf->kmax};
      lo[d3] = 0;
      hi[d3] = 1;
      /* non-reflecting inlet / outlet: average and maximum outward Mach number of the
       * boundary-adjacent cells of this patch (ref: src/procBlock.cpp:6235-6261) */
      orc_bc_extra ex;
      memset(&ex, 0, sizeof(ex));
      const aither_bc_state *sbc =
          (bcType == AITHER_BC_PRESSURE_OUTLET || bcType == AITHER_BC_INLET) ? bc_data(h, sf->tag)
                                                                               : NULL;
      const int nonRefl = sbc && sbc->isNonreflecting;
      if (nonRefl) {
        double avg = 0.0, mx = -DBL_MAX;
        long cnt = 0;
        for (int kk = lo[2]; kk < hi[2]; ++kk)
          for (int jj = lo[1]; jj < hi[1]; ++jj)
            for (int ii = lo[0]; ii < hi[0]; ++ii) {
              int ca[3] = {ii, jj, kk}, cf[3] = {ii, jj, kk};
              ca[d3] = src;
              cf[d3] = bnd;
              const double *fa =
                  d3 == 0 ? b->fAI + 4 * fidxI(b, cf[0], cf[1], cf[2])
                          : (d3 == 1 ? b->fAJ + 4 * fidxJ(b, cf[0], cf[1], cf[2])
                                     : b->fAK + 4 * fidxK(b, cf[0], cf[1], cf[2]));
              const double sg = st % 2 == 1 ? -1.0 : 1.0;
              const double *bs = b->state + neq * cidx(b, ca[0], ca[1], ca[2]);
              const double mach = (bs[h->ns] * (sg * fa[0]) + bs[h->ns + 1] * (sg * fa[1]) +
                                   bs[h->ns + 2] * (sg * fa[2])) / sos_of(h, bs);
              mx = mach > mx ? mach : mx;
              avg += mach;
              ++cnt;
            }
        ex.avgMach = avg / (double)cnt;
        ex.maxMach = mx;
      }
      for (int kk = lo[2]; kk < hi[2]; ++kk) {
        for (int jj = lo[1]; jj < hi[1]; ++jj) {
          for (int ii = lo[0]; ii < hi[0]; ++ii) {
            int ci[3] = {ii, jj, kk}, cg[3] = {ii, jj, kk}, cf[3] = {ii, jj, kk};
            ci[d3] = src;
            cg[d3] = gCell;
            cf[d3] = bnd;
            if (nonRefl) { /* ref: src/procBlock.cpp:2506-2517 */
              int ca[3] = {ii, jj, kk};
              ca[d3] = aCell;
              const long pa = pidx(b, ca[0], ca[1], ca[2]);
              ex.dt = b->dt[pa];
              cons_to_prim(h, b->consN + neq * pa, ex.stateN);
              for (int q = 0; q < 3; ++q) ex.pressGrad[q] = b->pressGrad[3 * pa + q];
              for (int q = 0; q < 9; ++q)
                ex.velGrad[q] = b->velGrad[9 * cidx(b, ca[0], ca[1], ca[2]) + q];
              g_bc_extra = &ex;
            }
            const double *fa =
                d3 == 0 ? b->fAI + 4 * fidxI(b, cf[0], cf[1], cf[2])
                        : (d3 == 1 ? b->fAJ + 4 * fidxJ(b, cf[0], cf[1], cf[2])
                                   : b->fAK + 4 * fidxK(b, cf[0], cf[1], cf[2]));
            double ghost[MAXEQ];
            ghost_state(h, b->state + neq * cidx(b, ci[0], ci[1], ci[2]), bcType,
                        fa, st, sf->tag, layer, 0.0, 0.0, ghost, NULL);
            g_bc_extra = NULL;
            memcpy(b->state + neq * cidx(b, cg[0], cg[1], cg[2]), ghost,
                   sizeof(double) * neq);
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* residual                                                                   */
static void face_states(const orc_level *h, const orc_block *b, int d, int i,
                        int j, int k, double *fl, double *fr) {
  /* face (i,j,k) of direction d lies between cells (..-1) and (..) */
  const int neq = h->neq;
  const int di = d == 0, dj = d == 1, dk = d == 2;
  const double *cw = d == 0 ? b->cwI : (d == 1 ? b->cwJ : b->cwK);
#define ST(o) (b->state + neq * cidx(b, i + (o)*di, j + (o)*dj, k + (o)*dk))
#define CW(o) (cw[cidx(b, i + (o)*di, j + (o)*dj, k + (o)*dk)])
  if (h->cfg.recon == AITHER_RECON_CONSTANT) {
    memcpy(fl, ST(-1), sizeof(double) * neq);
    memcpy(fr, ST(0), sizeof(double) * neq);
  } else if (h->cfg.recon == AITHER_RECON_MUSCL) {
    /* ref: src/procBlock.cpp:406-418 */
    orc_muscl(ST(-2), ST(-1), ST(0), neq, h->cfg.kappa, h->cfg.limiter, CW(-2),
              CW(-1), CW(0), fl);
    orc_muscl(ST(1), ST(0), ST(-1), neq, h->cfg.kappa, h->cfg.limiter, CW(1),
              CW(0), CW(-1), fr);
  } else {
    /* ref: src/procBlock.cpp:420-436 */
    const double *ul[5] = {ST(-3), ST(-2), ST(-1), ST(0), ST(1)};
    const double wl[5] = {CW(-3), CW(-2), CW(-1), CW(0), CW(1)};
    orc_weno(ul, wl, neq, h->cfg.recon == AITHER_RECON_WENOZ, fl);
    const double *ur[5] = {ST(2), ST(1), ST(0), ST(-1), ST(-2)};
    const double wr[5] = {CW(2), CW(1), CW(0), CW(-1), CW(-2)};
    orc_weno(ur, wr, neq, h->cfg.recon == AITHER_RECON_WENOZ, fr);
  }
#undef ST
#undef CW
}

/* ref: src/procBlock.cpp:384-491, :522-629, :660-767 (CalcInvFluxI/J/K) */
static void rusanov_flux_jacobian(const orc_level *h, const double *s,
                                  const double *area, int positive, double *J);
static double turb_inv_cell_spec_rad(const orc_level *h, const double *s,
                                     const double *fL, const double *fR);
static void calc_inv_flux(orc_level *h, orc_block *b, int d) {
  const int neq = h->neq;
  const int ni = b->ni + (d == 0), nj = b->nj + (d == 1), nk = b->nk + (d == 2);
  const int nd = d == 0 ? b->ni : (d == 1 ? b->nj : b->nk);
  const double *fA = d == 0 ? b->fAI : (d == 1 ? b->fAJ : b->fAK);
  for (int kk = 0; kk < nk; ++kk) {
    for (int jj = 0; jj < nj; ++jj) {
      for (int ii = 0; ii < ni; ++ii) {
        const int fi = d == 0 ? ii : (d == 1 ? jj : kk); /* face index along d */
        double fl[MAXEQ], fr[MAXEQ], flux[MAXEQ];
        face_states(h, b, d, ii, jj, kk, fl, fr);
        const long f0 = d == 0 ? fidxI(b, ii, jj, kk)
                               : (d == 1 ? fidxJ(b, ii, jj, kk)
                                         : fidxK(b, ii, jj, kk));
        const double *area = fA + 4 * f0;
        inviscid_flux(h, fl, fr, area, flux);
        if (fi > 0) {
          double *r = b->residual +
                      neq * pidx(b, ii - (d == 0), jj - (d == 1), kk - (d == 2));
          for (int e = 0; e < neq; ++e) r[e] += flux[e] * area[3];
          if (h->cfg.isBlockMatrix) { /* ref: src/procBlock.cpp:452-457 */
            double J[MAXJAC];
            rusanov_flux_jacobian(h, fl, area, 1, J);
            double *a = b->a + h->asz * pidx(b, ii - (d == 0), jj - (d == 1), kk - (d == 2));
            for (int q = 0; q < h->asz; ++q) a[q] += J[q];
          }
        }
        if (fi < nd) {
          double *r = b->residual + neq * pidx(b, ii, jj, kk);
          for (int e = 0; e < neq; ++e) r[e] -= flux[e] * area[3];
          const long f1 = d == 0 ? fidxI(b, ii + 1, jj, kk)
                                 : (d == 1 ? fidxJ(b, ii, jj + 1, kk)
                                           : fidxK(b, ii, jj, kk + 1));
          const double sr = inv_cell_spec_rad(
              h, b->state + neq * cidx(b, ii, jj, kk), area, fA + 4 * f1);
          double *sp = b->specRad + 2 * pidx(b, ii, jj, kk);
          /* turbModel::InviscidCellSpecRad; ref: src/procBlock.cpp:473-478 */
          const double tsr =
              h->nt > 0 ? turb_inv_cell_spec_rad(h, b->state + neq * cidx(b, ii, jj, kk),
                                                 area, fA + 4 * f1)
                        : 0.0;
          sp[0] += sr;
          sp[1] += tsr;
          if (!h->cfg.isBlockMatrix) {
            double *a = b->a + h->asz * pidx(b, ii, jj, kk);
            a[0] += sr;
            if (h->nt > 0) a[1] += tsr;
          } else { /* ref: src/procBlock.cpp:481-486 */
            double J[MAXJAC];
            rusanov_flux_jacobian(h, fr, area, 0, J);
            double *a = b->a + h->asz * pidx(b, ii, jj, kk);
            for (int q = 0; q < h->asz; ++q) a[q] -= J[q];
          }
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* transport: Sutherland (ref: src/transport.cpp:113-196); single species     */
static double species_viscosity(const orc_level *h, double t, int ss) {
  const double temp = t * h->cfg.tRef;
  const double mu =
      (h->cfg.suthViscC1[ss] * pow(temp, 1.5)) / (temp + h->cfg.suthViscS[ss]);
  return mu / h->cfg.muMixRef;
}
/* sutherland::MoleFractions; ref: src/transport.cpp:133-146 */
static void mole_fractions(const orc_level *h, const double *mf, double *x) {
  double sum = 0.0;
  for (int ii = 0; ii < h->ns; ++ii) {
    x[ii] = mf[ii] / h->cfg.molarMass[ii];
    sum += x[ii];
  }
  for (int ii = 0; ii < h->ns; ++ii) x[ii] /= sum;
}
/* sutherland::Viscosity with Wilke's mixing rule; ref: src/transport.cpp:70-96,148-162 */
static double viscosity_of(const orc_level *h, double t, const double *s) {
  if (h->ns == 1) return species_viscosity(h, t, 0);
  double mf[AITHER_MAX_SPECIES], x[AITHER_MAX_SPECIES], sv[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  mole_fractions(h, mf, x);
  for (int ii = 0; ii < h->ns; ++ii) sv[ii] = species_viscosity(h, t, ii);
  const double *mm = h->cfg.molarMass;
  double mixtureVisc = 0.0;
  for (int ii = 0; ii < h->ns; ++ii) {
    const double moleVisc = x[ii] * sv[ii];
    double denom = 0.0;
    for (int jj = 0; jj < h->ns; ++jj)
      denom += x[jj] / sqrt(1.0 + mm[ii] / mm[jj]) *
               pow(1.0 + sqrt(sv[ii] / sv[jj]) * pow(mm[jj] / mm[ii], 0.25), 2.0);
    mixtureVisc += moleVisc / denom;
  }
  return 4.0 / sqrt(2.0) * mixtureVisc;
}
static double species_conductivity(const orc_level *h, double t, int ss) {
  const double temp = t * h->cfg.tRef;
  const double k =
      (h->cfg.suthCondC1[ss] * pow(temp, 1.5)) / (temp + h->cfg.suthCondS[ss]);
  return k / h->cfg.kMixRef;
}
/* sutherland::Conductivity; ref: src/transport.cpp:97-110,178-192 (WilkesCond) */
static double conductivity_of(const orc_level *h, double t, const double *s) {
  if (h->ns == 1) return species_conductivity(h, t, 0);
  double mf[AITHER_MAX_SPECIES], x[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  mole_fractions(h, mf, x);
  double weightedAvg = 0.0, harmonicAvg = 0.0;
  for (int ii = 0; ii < h->ns; ++ii) {
    const double sc = species_conductivity(h, t, ii);
    weightedAvg += x[ii] * sc;
    harmonicAvg += x[ii] / sc;
  }
  harmonicAvg = 1.0 / harmonicAvg;
  return 0.5 * (weightedAvg + harmonicAvg);
}
static double eff_conductivity(const orc_level *h, double t, const double *s) {
  return conductivity_of(h, t, s) * h->cfg.nondimScaling;
}
/* ref: include/thermodynamic.hpp:61-64 */
static double prandtl_of(double gamma) {
  return (4.0 * gamma) / (9.0 * gamma - 5.0);
}

/* ref: src/procBlock.cpp:6171-6190 (UpdateAuxillaryVariables) */
static void update_aux(orc_level *h, orc_block *b) {
  for (int kk = -b->g; kk < b->nk + b->g; ++kk)
    for (int jj = -b->g; jj < b->nj + b->g; ++jj)
      for (int ii = -b->g; ii < b->ni + b->g; ++ii) {
        const int oi = ii < 0 || ii >= b->ni, oj = jj < 0 || jj >= b->nj,
                  ok = kk < 0 || kk >= b->nk;
        if (oi + oj + ok == 3) continue; /* corner */
        const long c = cidx(b, ii, jj, kk);
        b->temperature[c] = temperature_of(h, b->state + h->neq * c);
        if (h->cfg.isViscous)
          b->viscosity[c] = viscosity_of(h, b->temperature[c], b->state + h->neq * c);
      }
}

/* ------------------------------------------------------------------------ */
/* edge ghost cells and viscous-wall ghost cells                              */
/* multiArray3d::operator()(dir, d1, d2, d3): include/multiArray3d.hpp:231-243 */
static void dir_ijk(int dd, int d1, int d2, int d3, int *c) {
  if (dd == 0) { c[0] = d1; c[1] = d2; c[2] = d3; }
  else if (dd == 1) { c[0] = d3; c[1] = d1; c[2] = d2; }
  else { c[0] = d2; c[1] = d3; c[2] = d1; }
}
/* ref: src/boundaryConditions.cpp:109-185 (GetBCSurface) */
static const aither_surface *find_surface(const orc_block *b, int i, int j,
                                          int k, int surf) {
  for (int s = 0; s < b->nsurf; ++s) {
    const aither_surface *sf = &b->surf[s];
    const int st = surface_type(sf);
    if ((st - 1) / 2 != (surf - 1) / 2) continue;
    int in;
    if (surf <= 2)
      in = i >= sf->imin && i <= sf->imax && j >= sf->jmin && j < sf->jmax &&
           k >= sf->kmin && k < sf->kmax;
    else if (surf <= 4)
      in = i >= sf->imin && i < sf->imax && j >= sf->jmin && j <= sf->jmax &&
           k >= sf->kmin && k < sf->kmax;
    else
      in = i >= sf->imin && i < sf->imax && j >= sf->jmin && j < sf->jmax &&
           k >= sf->kmin && k <= sf->kmax;
    if (in) return sf;
  }
  fprintf(stderr, "oracle: no boundary surface at %d %d %d (surf %d)\n", i, j, k,
          surf);
  abort();
  return NULL;
}
static const double *face_area_dir(const orc_block *b, int faceDir, int dd,
                                   int d1, int d2, int d3) {
  int c[3];
  dir_ijk(dd, d1, d2, d3, c);
  return faceDir == 0 ? b->fAI + 4 * fidxI(b, c[0], c[1], c[2])
                      : (faceDir == 1 ? b->fAJ + 4 * fidxJ(b, c[0], c[1], c[2])
                                      : b->fAK + 4 * fidxK(b, c[0], c[1], c[2]));
}
/* ref: src/procBlock.cpp:2565-2703 (AssignInviscidGhostCellsEdge, viscous = 0)
 * and :2873-3025 (AssignViscousGhostCellsEdge, viscous = 1) */
static void assign_ghost_edges(orc_level *h, orc_block *b, int viscous) {
  const int neq = h->neq, g = b->g;
  const int nd[3] = {b->ni, b->nj, b->nk};
  for (int dd = 0; dd < 3; ++dd) {
    const int max1 = nd[dd], max2 = nd[(dd + 1) % 3], max3 = nd[(dd + 2) % 3];
    const int surfStart2 = 2 * ((dd + 1) % 3) + 1, surfStart3 = 2 * ((dd + 2) % 3) + 1;
    const int fd2 = (dd + 1) % 3, fd3 = (dd + 2) % 3; /* face arrays of dirs 2, 3 */
    for (int layer3 = 1; layer3 <= g; ++layer3)
      for (int layer2 = 1; layer2 <= g; ++layer2)
        for (int cc = 0; cc < 4; ++cc) {
          const int upper2 = cc > 1, upper3 = cc % 2 == 1;
          const int pCellD2 = upper2 ? max2 + layer2 - 2 : 1 - layer2;
          const int gCellD2 = upper2 ? pCellD2 + 1 : pCellD2 - 1;
          const int pCellD3 = upper3 ? max3 + layer3 - 2 : 1 - layer3;
          const int gCellD3 = upper3 ? pCellD3 + 1 : pCellD3 - 1;
          const int surf2 = upper2 ? surfStart2 + 1 : surfStart2;
          const int surf3 = upper3 ? surfStart3 + 1 : surfStart3;
          const int cFaceD2_2 = upper2 ? max2 : 0, cFaceD2_3 = upper3 ? max3 - 1 : 0;
          const int cFaceD3_2 = upper2 ? max2 - 1 : 0, cFaceD3_3 = upper3 ? max3 : 0;
          for (int d1 = 0; d1 < max1; ++d1) {
            int c2[3], c3[3], cg[3], cp2[3], cp3[3];
            dir_ijk(dd, d1, cFaceD2_2, cFaceD2_3, c2);
            dir_ijk(dd, d1, cFaceD3_2, cFaceD3_3, c3);
            const aither_surface *s2 = find_surface(b, c2[0], c2[1], c2[2], surf2);
            const aither_surface *s3 = find_surface(b, c3[0], c3[1], c3[2], surf3);
            const double *fArea2 = face_area_dir(b, fd2, dd, d1, cFaceD2_2, gCellD3);
            const double *fArea3 = face_area_dir(b, fd3, dd, d1, gCellD2, cFaceD3_3);
            int bc2 = s2->type, bc3 = s3->type;
            if (!viscous) {
              if (bc2 == AITHER_BC_VISCOUS_WALL) bc2 = AITHER_BC_SLIP_WALL;
              if (bc3 == AITHER_BC_VISCOUS_WALL) bc3 = AITHER_BC_SLIP_WALL;
            }
            int cw2[3], cw3[3];
            dir_ijk(dd, d1, cFaceD3_2, gCellD3, cw2);
            dir_ijk(dd, d1, gCellD2, cFaceD2_3, cw3);
            const double wDist2 = b->wallDist ? b->wallDist[cidx(b, cw2[0], cw2[1], cw2[2])] : 0.0;
            const double wDist3 = b->wallDist ? b->wallDist[cidx(b, cw3[0], cw3[1], cw3[2])] : 0.0;
            dir_ijk(dd, d1, gCellD2, gCellD3, cg);
            dir_ijk(dd, d1, pCellD2, gCellD3, cp2);
            dir_ijk(dd, d1, gCellD2, pCellD3, cp3);
            double *dst = b->state + neq * cidx(b, cg[0], cg[1], cg[2]);
            const double *from2 = b->state + neq * cidx(b, cp2[0], cp2[1], cp2[2]);
            const double *from3 = b->state + neq * cidx(b, cp3[0], cp3[1], cp3[2]);
            double ghost[MAXEQ];
            /* the viscous variant tests the literal string "slipWall" while the
             * types are unmapped there (:2994-3005), so in it a viscous wall is
             * never extended by this branch pair -- only the both-viscousWall
             * averaging below applies */
            if (bc2 == AITHER_BC_SLIP_WALL && bc3 != AITHER_BC_SLIP_WALL) {
              ghost_state(h, from2, bc2, fArea2, surf2, s2->tag, layer2, wDist2, 0.0, ghost, NULL);
              memcpy(dst, ghost, sizeof(double) * neq);
            } else if (bc2 != AITHER_BC_SLIP_WALL && bc3 == AITHER_BC_SLIP_WALL) {
              ghost_state(h, from3, bc3, fArea3, surf3, s3->tag, layer3, wDist3, 0.0, ghost, NULL);
              memcpy(dst, ghost, sizeof(double) * neq);
            } else if (!viscous || (bc2 == AITHER_BC_VISCOUS_WALL &&
                                    bc3 == AITHER_BC_VISCOUS_WALL)) {
              if (layer2 == layer3) {
                for (int e = 0; e < neq; ++e) ghost[e] = 0.5 * (from2[e] + from3[e]);
                memcpy(dst, ghost, sizeof(double) * neq);
              } else if (layer2 > layer3) {
                memcpy(ghost, from3, sizeof(double) * neq);
                memcpy(dst, ghost, sizeof(double) * neq);
              } else {
                memcpy(ghost, from2, sizeof(double) * neq);
                memcpy(dst, ghost, sizeof(double) * neq);
              }
            }
          }
        }
  }
}

/* ref: src/procBlock.cpp:2760-2836 (AssignViscousGhostCells) */
static void assign_viscous_ghosts(orc_level *h, orc_block *b) {
  const int neq = h->neq;
  for (int layer = 1; layer <= b->g; ++layer) {
    for (int s = 0; s < b->nsurf; ++s) {
      const aither_surface *sf = &b->surf[s];
      if (sf->type != AITHER_BC_VISCOUS_WALL) continue;
      const int st = surface_type(sf);
      const int d3 = (st - 1) / 2;
      const int nd[3] = {b->ni, b->nj, b->nk};
      const int r3 = d3 == 0 ? sf->imin : (d3 == 1 ? sf->jmin : sf->kmin);
      int gCell, iCell, aCell;
      if (st % 2 == 0) {
        gCell = r3 + layer - 1;
        iCell = r3 - layer;
        aCell = r3 - 1;
        if (iCell < 0) iCell = 0;
      } else {
        gCell = r3 - layer;
        iCell = r3 + layer - 1;
        aCell = r3;
        if (iCell >= nd[d3]) iCell = nd[d3] - 1;
      }
      int lo[3] = {sf->imin, sf->jmin, sf->kmin};
      int hi[3] = {sf->imax, sf->jmax, sf->kmax};
      lo[d3] = 0;
      hi[d3] = 1;
      for (int kk = lo[2]; kk < hi[2]; ++kk)
        for (int jj = lo[1]; jj < hi[1]; ++jj)
          for (int ii = lo[0]; ii < hi[0]; ++ii) {
            int ci[3] = {ii, jj, kk}, cg[3] = {ii, jj, kk}, cf[3] = {ii, jj, kk},
                ca[3] = {ii, jj, kk};
            ci[d3] = iCell;
            cg[d3] = gCell;
            cf[d3] = r3;
            ca[d3] = aCell;
            const double *fa =
                d3 == 0 ? b->fAI + 4 * fidxI(b, cf[0], cf[1], cf[2])
                        : (d3 == 1 ? b->fAJ + 4 * fidxJ(b, cf[0], cf[1], cf[2])
                                   : b->fAK + 4 * fidxK(b, cf[0], cf[1], cf[2]));
            const double wd = b->wallDist ? b->wallDist[cidx(b, ca[0], ca[1], ca[2])] : 0.0;
            /* nu at the wall-adjacent cell from the stored (previous evaluation's)
             * viscosity; ref: src/procBlock.cpp:2814-2822 */
            const long ac = cidx(b, ca[0], ca[1], ca[2]);
            const double nuW = b->viscosity[ac] / rho_of(h, b->state + neq * ac);
            double ghost[MAXEQ];
            orc_wall_vars wv;
            ghost_state(h, b->state + neq * cidx(b, ci[0], ci[1], ci[2]),
                        AITHER_BC_VISCOUS_WALL, fa, st, sf->tag, layer, wd, nuW, ghost, &wv);
            /* wallData_ keeps the first layer's record; ref: src/procBlock.cpp:6287-6290 */
            if (layer == 1)
              b->wall[s][(ii - lo[0]) + (long)(hi[0] - lo[0]) * ((jj - lo[1]) +
                         (long)(hi[1] - lo[1]) * (kk - lo[2]))] = wv;
            memcpy(b->state + neq * cidx(b, cg[0], cg[1], cg[2]), ghost,
                   sizeof(double) * neq);
          }
    }
  }
  assign_ghost_edges(h, b, 1);
}

/* ------------------------------------------------------------------------ */
/* viscous fluxes                                                             */
static void area_vec(const double *fa, double *v) {
  v[0] = fa[0] * fa[3]; /* unitVec3dMag::Vector(): unit * mag */
  v[1] = fa[1] * fa[3];
  v[2] = fa[2] * fa[3];
}
static const double *farea(const orc_block *b, int d, int i, int j, int k) {
  return d == 0 ? b->fAI + 4 * fidxI(b, i, j, k)
                : (d == 1 ? b->fAJ + 4 * fidxJ(b, i, j, k)
                          : b->fAK + 4 * fidxK(b, i, j, k));
}
/* Green-Gauss gradients of velocity and temperature on the control volume
 * centred on face (i,j,k) of direction d; ref: src/procBlock.cpp:5173-5303
 * (I), :5378-5508 (J), :5584-5714 (K); src/utility.cpp:59-175 */
static double g_face_press_grad[3]; /* pressure gradient of the last face_gradients call */
static void face_gradients(const orc_level *h, const orc_block *b, int d, int i,
                           int j, int k, double vg[9], double tg[3],
                           double kg[3], double wg[3], double (*mg)[3]) {
  const int ns = h->ns, neq = h->neq;
  const int e3[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const int *ed = e3[d];
  /* areas of the alternate control volume, per direction q: lower, upper */
  double al[3][3], au[3][3];
  for (int q = 0; q < 3; ++q) {
    double a0[3], a1[3];
    if (q == d) {
      area_vec(farea(b, d, i, j, k), a0);
      area_vec(farea(b, d, i + ed[0], j + ed[1], k + ed[2]), a1);
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      area_vec(farea(b, d, i - ed[0], j - ed[1], k - ed[2]), a1);
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    } else {
      const int *eq = e3[q];
      area_vec(farea(b, q, i + eq[0], j + eq[1], k + eq[2]), a0);
      area_vec(farea(b, q, i + eq[0] - ed[0], j + eq[1] - ed[1], k + eq[2] - ed[2]), a1);
      for (int c = 0; c < 3; ++c) au[q][c] = 0.5 * (a0[c] + a1[c]);
      area_vec(farea(b, q, i, j, k), a0);
      area_vec(farea(b, q, i - ed[0], j - ed[1], k - ed[2]), a1);
      for (int c = 0; c < 3; ++c) al[q][c] = 0.5 * (a0[c] + a1[c]);
    }
  }
  const double vol = 0.5 * (b->vol[cidx(b, i - ed[0], j - ed[1], k - ed[2])] +
                            b->vol[cidx(b, i, j, k)]);
  /* values on the faces of the control volume: 3 velocity components + T
   * (+ k, omega for RANS; ref: src/procBlock.cpp:5305-5352) */
  double vl[3][6], vu[3][6];
  const int nval = h->nt > 0 ? 6 : 4;
  const long cLo = cidx(b, i - ed[0], j - ed[1], k - ed[2]), cHi = cidx(b, i, j, k);
  for (int c = 0; c < nval; ++c) {
#define VAL(cell)                                                       \
  (c < 3 ? b->state[neq * (cell) + ns + c]                              \
         : (c == 3 ? b->temperature[(cell)] : b->state[neq * (cell) + ns + c]))
    for (int q = 0; q < 3; ++q) {
      if (q == d) {
        vl[q][c] = VAL(cLo);
        vu[q][c] = VAL(cHi);
      } else {
        const int *eq = e3[q];
        const long up0 = cidx(b, i + eq[0], j + eq[1], k + eq[2]);
        const long up1 = cidx(b, i + eq[0] - ed[0], j + eq[1] - ed[1], k + eq[2] - ed[2]);
        const long lo0 = cidx(b, i - eq[0], j - eq[1], k - eq[2]);
        const long lo1 = cidx(b, i - eq[0] - ed[0], j - eq[1] - ed[1], k - eq[2] - ed[2]);
        vu[q][c] = 0.25 * (VAL(cLo) + VAL(cHi) + VAL(up0) + VAL(up1));
        vl[q][c] = 0.25 * (VAL(cLo) + VAL(cHi) + VAL(lo0) + VAL(lo1));
      }
    }
#undef VAL
  }
  const double invVol = 1.0 / vol;
  /* velGrad(r, c) = d u_c / d x_r */
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double t = vu[0][c] * au[0][r] - vl[0][c] * al[0][r] +
                       vu[1][c] * au[1][r] - vl[1][c] * al[1][r] +
                       vu[2][c] * au[2][r] - vl[2][c] * al[2][r];
      vg[3 * r + c] = t * invVol;
    }
  for (int r = 0; r < 3; ++r) {
    const double t = vu[0][3] * au[0][r] - vl[0][3] * al[0][r] +
                     vu[1][3] * au[1][r] - vl[1][3] * al[1][r] +
                     vu[2][3] * au[2][r] - vl[2][3] * al[2][r];
    tg[r] = t * invVol;
  }
  { /* pressure gradient of the same control volume (ref: src/procBlock.cpp:5319-5331;
     * kept for the non-reflecting boundary conditions) */
    double pl[3], pu[3];
#define PV(cell) (b->state[neq * (cell) + ns + 3])
    for (int q = 0; q < 3; ++q) {
      if (q == d) {
        pl[q] = PV(cLo);
        pu[q] = PV(cHi);
      } else {
        const int *eq = e3[q];
        const long up0 = cidx(b, i + eq[0], j + eq[1], k + eq[2]);
        const long up1 = cidx(b, i + eq[0] - ed[0], j + eq[1] - ed[1], k + eq[2] - ed[2]);
        const long lo0 = cidx(b, i - eq[0], j - eq[1], k - eq[2]);
        const long lo1 = cidx(b, i - eq[0] - ed[0], j - eq[1] - ed[1], k - eq[2] - ed[2]);
        pu[q] = 0.25 * (PV(cLo) + PV(cHi) + PV(up0) + PV(up1));
        pl[q] = 0.25 * (PV(cLo) + PV(cHi) + PV(lo0) + PV(lo1));
      }
    }
#undef PV
    for (int r = 0; r < 3; ++r) {
      const double t = pu[0] * au[0][r] - pl[0] * al[0][r] + pu[1] * au[1][r] - pl[1] * al[1][r] +
                       pu[2] * au[2][r] - pl[2] * al[2][r];
      g_face_press_grad[r] = t * invVol;
    }
  }
  for (int c = 4; c < nval; ++c) {
    double *out = c == 4 ? kg : wg;
    for (int r = 0; r < 3; ++r) {
      const double t = vu[0][c] * au[0][r] - vl[0][c] * al[0][r] +
                       vu[1][c] * au[1][r] - vl[1][c] * al[1][r] +
                       vu[2][c] * au[2][r] - vl[2][c] * al[2][r];
      out[r] = t * invVol;
    }
  }
  if (nval == 4) {
    for (int r = 0; r < 3; ++r) kg[r] = wg[r] = 0.0;
  }
  /* mass-fraction gradients (multi-species; ref: src/procBlock.cpp:5347-5376) */
  if (ns > 1 && mg) {
    for (int ss = 0; ss < ns; ++ss) {
      double ml[3], mu_[3];
#define MFV(cell) (b->state[neq * (cell) + ss] / rho_of(h, b->state + neq * (cell)))
      for (int q = 0; q < 3; ++q) {
        if (q == d) {
          ml[q] = MFV(cLo);
          mu_[q] = MFV(cHi);
        } else {
          const int *eq = e3[q];
          const long up0 = cidx(b, i + eq[0], j + eq[1], k + eq[2]);
          const long up1 = cidx(b, i + eq[0] - ed[0], j + eq[1] - ed[1], k + eq[2] - ed[2]);
          const long lo0 = cidx(b, i - eq[0], j - eq[1], k - eq[2]);
          const long lo1 = cidx(b, i - eq[0] - ed[0], j - eq[1] - ed[1], k - eq[2] - ed[2]);
          mu_[q] = 0.25 * (MFV(cLo) + MFV(cHi) + MFV(up0) + MFV(up1));
          ml[q] = 0.25 * (MFV(cLo) + MFV(cHi) + MFV(lo0) + MFV(lo1));
        }
      }
#undef MFV
      for (int r = 0; r < 3; ++r) {
        const double t = mu_[0] * au[0][r] - ml[0] * al[0][r] + mu_[1] * au[1][r] -
                         ml[1] * al[1][r] + mu_[2] * au[2][r] - ml[2] * al[2][r];
        mg[ss][r] = t * invVol;
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* turbulence models: k-omega Wilcox 2006 and Menter SST 2003                 */
/* constants: include/turbulence.hpp:391-398 (Wilcox), :489-501 (SST) */
#define KW_GAMMA 0.52
#define KW_BETASTAR 0.09
#define KW_SIGMA 0.5
#define KW_SIGMASTAR 0.6
#define KW_SIGMAD0 0.125
#define KW_BETA0 0.0708
#define KW_CLIM 0.875
#define SST_BETASTAR 0.09
#define SST_SIGMAK1 0.85
#define SST_SIGMAK2 1.0
#define SST_SIGMAW1 0.5
#define SST_SIGMAW2 0.856
#define SST_BETA1 0.075
#define SST_BETA2 0.0828
#define SST_GAMMA1 (5.0 / 9.0)
#define SST_GAMMA2 0.44
#define SST_A1 0.31
#define SST_KPROD2DEST 10.0
static int is_sst(const orc_level *h) { return h->cfg.turbModel == AITHER_TURB_SST; }
static int is_rans(const orc_level *h) { return h->nt > 0; }
/* ref: include/turbulence.hpp:70 (0.9), :462 (Wilcox 8/9), :578 (SST 0.9) */
static double turb_prandtl(const orc_level *h) {
  return h->cfg.turbModel == AITHER_TURB_KW_WILCOX ? 8.0 / 9.0 : 0.9;
}
static double blended(double c1, double c2, double f1) { /* turbulence.cpp:592-595 */
  return f1 * c1 + (1.0 - f1) * c2;
}
static double sigma_k(const orc_level *h, double f1) { /* turbulence.hpp:476,599 */
  return is_sst(h) ? blended(SST_SIGMAK1, SST_SIGMAK2, f1) : KW_SIGMASTAR;
}
static double sigma_w(const orc_level *h, double f1) { /* turbulence.hpp:477,602 */
  return is_sst(h) ? blended(SST_SIGMAW1, SST_SIGMAW2, f1) : KW_SIGMA;
}
static double wall_beta(const orc_level *h) { /* turbulence.hpp:463,577 */
  return is_sst(h) ? SST_BETA1 : KW_BETA0;
}
static double tke_of(const orc_level *h, const double *s) { return s[h->ns + 4]; }
static double omega_of(const orc_level *h, const double *s) { return s[h->ns + 5]; }
/* ref: src/turbulence.cpp:38-40 */
static double eddy_visc_no_lim(const orc_level *h, const double *s) {
  return rho_of(h, s) * tke_of(h, s) / omega_of(h, s);
}
/* primitive::LimitTurb; ref: src/primitive.cpp:100-106, turbulence.hpp:72-73 */
/* ------------------------------------------------------------------------ */
/* wall law (Nichols & Nelson 2004 with White & Christoph's compressible law
 * of the wall); ref: src/wallLaw.cpp:31-289, include/wallLaw.hpp:36-95.
 * mode 0 adiabatic, 1 constant heat flux, 2 isothermal. The members of the
 * reference's wallLaw object after its last function evaluation are what the
 * wall variables are built from, so the context keeps them the same way.   */
typedef struct {
  const orc_level *h;
  const double *state;
  double mf[AITHER_MAX_SPECIES];
  double wallDist, vonKarmen, yplus0, velTanMag, tInt, cp;
  double beta, gamma, q, phi, yplusWhite, uStar, uplus, tW, rhoW, muW, mutW, kW, recovery;
  int mode;
  double heatFlux, yplus, temperature;
} wl_ctx;
static void wl_set_wall_vars(wl_ctx *c, double tW) { /* ref: src/wallLaw.cpp:229-237 */
  const orc_level *h = c->h;
  c->tW = tW;
  double R = 0.0; /* eos DensityTP, src/eos.cpp:111-115 */
  for (int ss = 0; ss < h->ns; ++ss) R += c->mf[ss] * h->cfg.gasConstant[ss];
  c->rhoW = c->state[h->ns + 3] / (R * tW);
  c->muW = viscosity_of(h, tW, c->state) * h->cfg.nondimScaling;
  c->kW = eff_conductivity(h, tW, c->state);
}
static double wl_func(wl_ctx *c, double yplus) {
  const orc_level *h = c->h;
  /* CalcVelocities; ref: :264-268 */
  c->uplus = (c->wallDist * c->rhoW * c->velTanMag) / (c->muW * yplus);
  c->uStar = c->velTanMag / c->uplus;
  if (c->mode == 1) { /* CalcWallTemperature; ref: :220-227, then SetWallVars */
    c->temperature = c->tInt + c->recovery * c->uStar * c->uStar * c->uplus * c->uplus /
                                   (2.0 * c->cp + c->heatFlux * c->muW / (c->rhoW * c->kW * c->uStar));
    wl_set_wall_vars(c, c->temperature);
  }
  /* UpdateGamma; ref: :186-191 */
  c->gamma = c->recovery * c->uStar * c->uStar / (2.0 * c->cp * c->tW);
  if (c->mode == 2) { /* CalcHeatFlux; ref: :211-218 */
    const double tmp = (c->tInt / c->tW - 1.0 + c->gamma * c->uplus * c->uplus) / c->uplus;
    c->heatFlux = tmp * (c->rhoW * c->tW * c->kW * c->uStar) / c->muW;
  }
  /* UpdateConstants; ref: :193-198 */
  c->beta = c->heatFlux * c->muW / (c->rhoW * c->tW * c->kW * c->uStar);
  c->q = sqrt(c->beta * c->beta + 4.0 * c->gamma);
  c->phi = asin(-c->beta / c->q);
  /* CalcYplusWhite; ref: :200-205 */
  c->yplusWhite = exp((c->vonKarmen / sqrt(c->gamma)) *
                      (asin((2.0 * c->gamma * c->uplus - c->beta) / c->q) - c->phi)) *
                  c->yplus0;
  c->yplus = yplus;
  /* CalcYplusRoot; ref: :239-244 */
  const double sixth = 1.0 / 6.0;
  const double ku = c->vonKarmen * c->uplus;
  (void)h;
  return yplus - (c->uplus + c->yplusWhite -
                  c->yplus0 * (1.0 + ku + 0.5 * ku * ku + sixth * pow(ku, 3.0)));
}
/* Ridder's method; ref: include/utility.hpp:130-184 */
static double wl_find_root(wl_ctx *c, double x1, double x2, double tol) {
  double f1 = wl_func(c, x1);
  double f2 = wl_func(c, x2);
  if (sign_of(f1) == sign_of(f2) && sign_of(f1) != 0.0) return 0.5 * (x1 + x2);
  double x4 = x1;
  for (int ii = 0; ii < 100; ++ii) {
    const double x3 = 0.5 * (x1 + x2);
    const double f3 = wl_func(c, x3);
    if (f3 == 0.0) return x3;
    const double denom = sqrt(fabs(f3 * f3 - f1 * f2));
    if (denom == 0.0) return x3;
    const double fac = sign_of(f1 - f2);
    x4 = x3 + (x3 - x1) * (fac * f3) / denom;
    const double f4 = wl_func(c, x4);
    if (f4 == 0.0) return x4;
    if (sign_of(f4) != sign_of(f3)) {
      x1 = x3;
      f1 = f3;
      x2 = x4;
      f2 = f4;
    } else if (sign_of(f4) != sign_of(f1)) {
      x2 = x4;
      f2 = f4;
    } else {
      x1 = x4;
      f1 = f4;
    }
    if (fabs(x2 - x1) <= tol) return x4;
  }
  return x4;
}
static void wall_law_eval(const orc_level *h, const aither_bc_state *bc, int mode,
                          const double *state, double wallDist, const double area[3],
                          int isLower, orc_wall_vars *wv) {
  const int ns = h->ns;
  wl_ctx c;
  memset(&c, 0, sizeof(c));
  c.h = h;
  c.state = state;
  c.mode = mode;
  c.wallDist = wallDist;
  c.vonKarmen = bc->vonKarmen;
  c.yplus0 = exp(-bc->vonKarmen * bc->wallConstant);
  mass_fractions(h, state, c.mf);
  /* tangential velocity relative to the wall; ref: :41-44 */
  double vel[3], velTan[3];
  for (int d = 0; d < 3; ++d) vel[d] = state[ns + d] - bc->velocity[d];
  const double vn = vel[0] * area[0] + vel[1] * area[1] + vel[2] * area[2];
  for (int d = 0; d < 3; ++d) velTan[d] = vel[d] - vn * area[d];
  c.velTanMag = sqrt(velTan[0] * velTan[0] + velTan[1] * velTan[1] + velTan[2] * velTan[2]);
  c.tInt = temperature_of(h, state);
  c.cp = cp_mix(h, c.mf);
  /* CalcRecoveryFactor; ref: :286-289 */
  c.recovery = pow(prandtl_of(c.cp / cv_mix(h, c.mf)), 1.0 / 3.0);
  if (mode == 0) { /* Crocco-Busemann wall temperature; ref: :46-52 */
    const double tW = c.tInt + 0.5 * c.recovery * c.velTanMag * c.velTanMag / c.cp;
    wl_set_wall_vars(&c, tW);
    c.heatFlux = 0.0;
  } else if (mode == 1) { /* ref: :96-105 */
    c.heatFlux = bc->heatFlux;
    c.temperature = c.tInt;
    wl_set_wall_vars(&c, c.tInt);
  } else { /* ref: :148-160 */
    c.temperature = bc->temperature;
    wl_set_wall_vars(&c, bc->temperature);
  }
  wl_find_root(&c, 1.0e1, 1.0e4, 1.0e-8);
  wv->tke = 0.0;
  wv->sdr = 0.0;
  if (h->nt > 0) { /* CalcTurbVars, EddyVisc; ref: :246-262, :270-284 */
    const double dYplusWhite =
        2.0 * c.yplusWhite * c.vonKarmen * sqrt(c.gamma) / c.q *
        sqrt(fmax(1.0 - pow(2.0 * c.gamma * c.uplus - c.beta, 2.0) / (c.q * c.q), 0.0));
    const double ku = c.vonKarmen * c.uplus;
    c.mutW = c.muW * (1.0 + dYplusWhite - c.vonKarmen * c.yplus0 * (1.0 + ku + 0.5 * ku * ku)) -
             viscosity_of(h, c.tInt, state) * h->cfg.nondimScaling;
    c.mutW = fmax(c.mutW, 0.0);
    double wi = 6.0 * c.muW / (wall_beta(h) * c.rhoW * wallDist * wallDist);
    wi *= h->cfg.nondimScaling;
    double wo = c.uStar / (sqrt(0.09) * c.vonKarmen * wallDist);
    wo *= h->cfg.nondimScaling;
    wv->sdr = sqrt(wi * wi + wo * wo);
    wv->tke = wv->sdr * c.mutW / rho_of(h, state) * (1.0 / h->cfg.nondimScaling);
  }
  wv->heatFlux = c.heatFlux;
  wv->yplus = c.yplus;
  wv->density = c.rhoW;
  wv->temperature = mode == 0 ? c.tW : c.temperature;
  wv->viscosity = c.muW;
  wv->turbEddyVisc = c.mutW;
  wv->frictionVelocity = c.uStar;
  const double tauMag = c.uStar * c.uStar * c.rhoW; /* ShearStressMag */
  for (int d = 0; d < 3; ++d) {
    wv->shearStress[d] = tauMag * velTan[d] / c.velTanMag;
    if (!isLower) wv->shearStress[d] *= -1.0;
  }
}

static void limit_turb(const orc_level *h, double *s) {
  for (int tt = 0; tt < h->nt; ++tt)
    s[h->ns + 4 + tt] = s[h->ns + 4 + tt] > 1.0e-20 ? s[h->ns + 4 + tt] : 1.0e-20;
}
static double ddot_trans(const double a[9], const double b[9]) {
  /* tensor::DoubleDotTrans: sum of the elementwise product, flat order;
   * ref: include/tensor.hpp:353-356 */
  double sum = 0.0;
  for (int q = 0; q < 9; ++q) sum += a[q] * b[q];
  return sum;
}
static double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
/* SST cross diffusion; ref: include/turbulence.hpp:528-537 */
static double sst_cdkw(const orc_level *h, const double *s, const double kg[3],
                       const double wg[3]) {
  const double v = 2.0 * rho_of(h, s) * SST_SIGMAW2 / omega_of(h, s) * dot3(kg, wg);
  return v > 1.0e-10 ? v : 1.0e-10;
}
/* turbModel::EddyViscAndBlending; ref: src/turbulence.cpp:405-423 (Wilcox,
 * OmegaTilda :329-342), :663-684 (SST, EddyVisc :570-580, F1/F2 :582-590,
 * Alpha1-3 :597-615) */
static void eddy_visc_and_blending(const orc_level *h, const double *s,
                                   const double vg[9], const double kg[3],
                                   const double wg[3], double mu, double wallDist,
                                   double *mut, double *f1, double *f2) {
  const double scaling = h->cfg.nondimScaling;
  const double rho = rho_of(h, s), tke = tke_of(h, s), omg = omega_of(h, s);
  const double trace = vg[0] + vg[4] + vg[8];
  if (!is_sst(h)) {
    *f1 = 1.0;
    *f2 = 0.0;
    double sHat[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        sHat[3 * r + c] = 0.5 * (vg[3 * r + c] + vg[3 * c + r]) -
                          1.0 / 3.0 * trace * (r == c ? 1.0 : 0.0);
    const double lim = scaling * KW_CLIM * sqrt(2.0 * ddot_trans(sHat, sHat) / KW_BETASTAR);
    const double omegaTilda = omg > lim ? omg : lim;
    *mut = rho * tke / omegaTilda;
    return;
  }
  const double dE = wallDist + ORC_EPS;
  const double alpha1 = scaling * sqrt(tke) / (SST_BETASTAR * omg * dE);
  const double alpha2 = scaling * scaling * 500.0 * mu / (dE * dE * rho * omg);
  const double cdkw = sst_cdkw(h, s, kg, wg);
  const double alpha3 = 4.0 * rho * SST_SIGMAW2 * tke / (cdkw * dE * dE);
  const double m12 = alpha1 > alpha2 ? alpha1 : alpha2;
  const double arg1 = m12 < alpha3 ? m12 : alpha3;
  *f1 = tanh(pow(arg1, 4.0));
  const double arg2 = 2.0 * alpha1 > alpha2 ? 2.0 * alpha1 : alpha2;
  *f2 = tanh(arg2 * arg2);
  double sr[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) sr[3 * r + c] = 0.5 * (vg[3 * r + c] + vg[3 * c + r]);
  const double meanStrainRate = sqrt(2.0 * ddot_trans(sr, sr));
  const double d1 = SST_A1 * omg, d2 = scaling * meanStrainRate * (*f2);
  *mut = rho * SST_A1 * tke / (d1 > d2 ? d1 : d2);
}
/* turbulence-equation spectral radii; ref: src/turbulence.cpp:152-171
 * (inviscid), :500-527 (Wilcox viscous: unlimited eddy viscosity), :783-808 (SST) */
static double turb_inv_cell_spec_rad(const orc_level *h, const double *s,
                                     const double *fL, const double *fR) {
  double na[3];
  for (int q = 0; q < 3; ++q) na[q] = 0.5 * (fL[q] + fR[q]);
  const double mag = sqrt(na[0] * na[0] + na[1] * na[1] + na[2] * na[2]);
  for (int q = 0; q < 3; ++q) na[q] /= mag;
  const double fMag = 0.5 * (fL[3] + fR[3]);
  return fabs(dot3(s + h->ns, na)) * fMag;
}
static double turb_inv_face_spec_rad(const orc_level *h, const double *s,
                                     const double *fA, int positive) {
  const double velNorm = dot3(s + h->ns, fA);
  return positive ? 0.5 * fA[3] * fabs(velNorm + fabs(velNorm))
                  : 0.5 * fA[3] * fabs(velNorm - fabs(velNorm));
}
static double turb_visc_spec_rad(const orc_level *h, const double *s, double length,
                                 double mu, double mut, double f1) {
  const double mt = is_sst(h) ? mut : eddy_visc_no_lim(h, s);
  return h->cfg.nondimScaling * length / rho_of(h, s) * (mu + sigma_k(h, f1) * mt);
}
/* turbModel::SrcSpecRad; ref: src/turbulence.cpp:438-443, :699-704 */
static double turb_src_spec_rad(const orc_level *h, const double *s, double vol) {
  const double invScaling = 1.0 / h->cfg.nondimScaling;
  return -2.0 * KW_BETASTAR * omega_of(h, s) * vol * invScaling;
}
/* turbModel::CalcTurbSrc; ref: src/turbulence.cpp:344-384 (Wilcox; Beta/FBeta/Xw
 * :291-319), :617-661 (SST); BoussinesqReynoldsStress :55-70 */
static void calc_turb_src(const orc_level *h, const double *s, const double vg[9],
                          const double kg[3], const double wg[3], double mut,
                          double f1, double src[2], double *betaOut) {
  const double scaling = h->cfg.nondimScaling, invScaling = 1.0 / scaling;
  const double rho = rho_of(h, s), tke = tke_of(h, s), omg = omega_of(h, s);
  const double trace = vg[0] + vg[4] + vg[8];
  /* tau = lambda tr(G) I + mut (G + G^T) - 2/3 rho k I; lambda = -2/3 mut */
  const double lambda = 0.0 - (2.0 / 3.0) * mut;
  double tau[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double I = r == c ? 1.0 : 0.0;
      tau[3 * r + c] = lambda * trace * I + mut * (vg[3 * r + c] + vg[3 * c + r]) -
                       2.0 / 3.0 * rho * tke * I;
    }
  const double prodRaw = scaling * ddot_trans(tau, vg);
  if (!is_sst(h)) {
    const double tkeDest = invScaling * KW_BETASTAR * (rho * tke * omg * 1.0);
    /* Xw = |Omega_ij Omega_jk S^_ki / (betaStar w)^3| scaling^3 */
    double vort[9], ski[9], vv[9];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        vort[3 * r + c] = 0.5 * (vg[3 * r + c] - vg[3 * c + r]);
        ski[3 * r + c] =
            0.5 * (vg[3 * r + c] + vg[3 * c + r] - trace * (r == c ? 1.0 : 0.0));
      }
    for (int q = 0; q < 9; ++q) vv[q] = 0.0;
    for (int cc = 0; cc < 3; ++cc) /* tensor::MatMult, include/tensor.hpp:272-282 */
      for (int rr = 0; rr < 3; ++rr)
        for (int ii = 0; ii < 3; ++ii) vv[3 * rr + ii] += vort[3 * rr + cc] * vort[3 * cc + ii];
    const double xw = fabs(ddot_trans(vv, ski) / pow(KW_BETASTAR * omg, 3.0)) *
                      pow(scaling, 3.0);
    const double beta = KW_BETA0 * ((1.0 + 85.0 * xw) / (1.0 + 100.0 * xw));
    *betaOut = beta;
    const double omgDest = invScaling * beta * (rho * omg * omg);
    double tkeProd = prodRaw > 0.0 ? prodRaw : 0.0;
    double omgProd = KW_GAMMA * omg / tke * tkeProd;
    omgProd = omgProd > 0.0 ? omgProd : 0.0;
    const double kw = dot3(kg, wg);
    const double sigmaD = kw <= 0.0 ? 0.0 : KW_SIGMAD0;
    const double omgCd = scaling * sigmaD * (rho / omg * kw);
    src[0] = tkeProd - tkeDest;
    src[1] = omgProd - omgDest + omgCd;
    return;
  }
  const double cdkw = sst_cdkw(h, s, kg, wg);
  const double gamma = blended(SST_GAMMA1, SST_GAMMA2, f1);
  const double beta = blended(SST_BETA1, SST_BETA2, f1);
  *betaOut = beta;
  const double tkeDest = invScaling * SST_BETASTAR * (rho * tke * omg * 1.0);
  const double omgDest = invScaling * beta * (rho * omg * omg);
  const double lim = SST_KPROD2DEST * tkeDest;
  double tkeProd = prodRaw < lim ? prodRaw : lim;
  tkeProd = tkeProd > 0.0 ? tkeProd : 0.0;
  double omgProd = gamma * rho / mut * tkeProd;
  omgProd = omgProd > 0.0 ? omgProd : 0.0;
  const double omgCd = scaling * (1.0 - f1) * cdkw;
  src[0] = tkeProd - tkeDest;
  src[1] = omgProd - omgDest + omgCd;
}
/* ref: src/utility.cpp:425-437 (TauNormal), src/transport.cpp:173-176 (Lambda) */
static void tau_normal(const double vg[9], const double n[3], double mu,
                       double mut, double tau[3]) {
  const double lambda = 0.0 - (2.0 / 3.0) * (mu + mut);
  const double trace = vg[0] + vg[4] + vg[8];
  for (int r = 0; r < 3; ++r) {
    double mm = 0.0; /* ((G + G^T) n)_r */
    for (int c = 0; c < 3; ++c) mm += (vg[3 * r + c] + vg[3 * c + r]) * n[c];
    tau[r] = lambda * trace * n[r] + (mu + mut) * mm;
  }
}
static double proj_c2c_dist(const orc_block *b, int d, int ii, int jj, int kk);
static void approx_tsl_jacobian(const orc_level *h, const double *s, double lamVisc,
                                double turbVisc, double f1, const double *area,
                                double dist, int left, const double vGrad[9],
                                double *J);
/* ref: src/procBlock.cpp:1233-1497 (CalcViscFluxI; J, K alike); low-Re walls */
static void calc_visc_flux(orc_level *h, orc_block *b, int d) {
  const int neq = h->neq, ns = h->ns;
  const int nd[3] = {b->ni, b->nj, b->nk};
  const int di = d == 0, dj = d == 1, dk = d == 2;
  const double *cw = d == 0 ? b->cwI : (d == 1 ? b->cwJ : b->cwK);
  const double viscCoeff = h->cfg.viscousCFLCoeff;
  const double sixth = 1.0 / 6.0;
  const int rans = is_rans(h);
  for (int kk = 0; kk < b->nk + dk; ++kk)
    for (int jj = 0; jj < b->nj + dj; ++jj)
      for (int ii = 0; ii < b->ni + di; ++ii) {
        const int fi = d == 0 ? ii : (d == 1 ? jj : kk);
        double vg[9], tg[3], kg[3], wg[3], mg[AITHER_MAX_SPECIES][3];
        face_gradients(h, b, d, ii, jj, kk, vg, tg, kg, wg, mg);
        const double pg[3] = {g_face_press_grad[0], g_face_press_grad[1], g_face_press_grad[2]};
        double state[MAXEQ], mu, wDist = 0.0;
#define CL(o) cidx(b, ii + (o)*di, jj + (o)*dj, kk + (o)*dk)
        double mut = 0.0, f1 = 0.0, f2 = 0.0, flux[MAXEQ];
        const double *fa = farea(b, d, ii, jj, kk);
        for (int e = 0; e < neq; ++e) flux[e] = 0.0;
        /* boundary face on a viscous wall: low-Re treatment, or the wall law where its
         * y+ stayed >= 10 (ref: src/procBlock.cpp:1269-1300) */
        int isLowReWall = 0, isWallLawFace = 0;
        const orc_wall_vars *wvp = NULL;
        const aither_bc_state *wbc = NULL;
        if (fi == 0 || fi == nd[d]) {
          const aither_surface *sf = find_surface(b, ii, jj, kk, 2 * d + (fi == 0 ? 1 : 2));
          if (sf->type == AITHER_BC_VISCOUS_WALL) {
            wbc = bc_data(h, sf->tag);
            if (wbc && wbc->isWallLaw) {
              int lo[3] = {sf->imin, sf->jmin, sf->kmin}, hi[3] = {sf->imax, sf->jmax, sf->kmax};
              int c3[3] = {ii, jj, kk};
              lo[d] = 0;
              hi[d] = 1;
              c3[d] = 0;
              wvp = &b->wall[sf - b->surf][(c3[0] - lo[0]) + (long)(hi[0] - lo[0]) *
                                           ((c3[1] - lo[1]) + (long)(hi[1] - lo[1]) * (c3[2] - lo[2]))];
              isWallLawFace = !(wvp->yplus < 10.0);
            }
            isLowReWall = !isWallLawFace;
          }
        }
        if (isWallLawFace) {
          /* wall state, wall viscosities, prescribed stress and heat flux
           * (ref: src/procBlock.cpp:1286-1300, src/wallData.cpp:299-313,
           * viscousFlux::CalcWallLawFlux src/viscousFlux.cpp:213-249) */
          const double invScaling = 1.0 / h->cfg.nondimScaling;
          f1 = 1.0;
          f2 = 1.0;
          mu = wvp->viscosity * invScaling;
          mut = wvp->turbEddyVisc * invScaling;
          double mfw[AITHER_MAX_SPECIES], pW = 0.0;
          /* wall mass fractions are the adjacent cell's (ref: src/ghostStates.cpp:146) */
          mass_fractions(h, b->state + neq * (fi == 0 ? CL(0) : CL(-1)), mfw);
          for (int ss = 0; ss < ns; ++ss) {
            state[ss] = mfw[ss] * wvp->density;
            pW += state[ss] * h->cfg.gasConstant[ss]; /* eos PressureRT */
          }
          state[ns] = wbc->velocity[0];
          state[ns + 1] = wbc->velocity[1];
          state[ns + 2] = wbc->velocity[2];
          state[ns + 3] = pW * wvp->temperature;
          if (rans) {
            state[ns + 4] = wvp->tke;
            state[ns + 5] = wvp->sdr;
          }
          flux[ns] = wvp->shearStress[0];
          flux[ns + 1] = wvp->shearStress[1];
          flux[ns + 2] = wvp->shearStress[2];
          flux[ns + 3] = (wvp->shearStress[0] * wbc->velocity[0] + wvp->shearStress[1] * wbc->velocity[1] +
                          wvp->shearStress[2] * wbc->velocity[2]) + wvp->heatFlux;
          if (rans) { /* WallSigmaK / WallSigmaW: include/turbulence.hpp:478-479, :605-606 */
            const double sk = is_sst(h) ? 0.85 : 0.6, sw = 0.5;
            flux[ns + 4] = (wvp->viscosity + sk * wvp->turbEddyVisc) * dot3(kg, fa);
            flux[ns + 5] = (wvp->viscosity + sw * wvp->turbEddyVisc) * dot3(wg, fa);
          }
        } else {
        if (h->cfg.viscRecon == 0) { /* central; ref: :1305-1321 */
          const double w[2] = {cw[CL(-1)], cw[CL(0)]};
          double c[2];
          lagrange_coeff(w, 1, 0, 0, c);
          for (int e = 0; e < neq; ++e)
            state[e] = c[0] * b->state[neq * CL(0) + e] + c[1] * b->state[neq * CL(-1) + e];
          mu = c[0] * b->viscosity[CL(0)] + c[1] * b->viscosity[CL(-1)];
          if (b->wallDist) wDist = c[0] * b->wallDist[CL(0)] + c[1] * b->wallDist[CL(-1)];
        } else { /* centralFourth; ref: :1323-1346, include/reconstruction.hpp:359-379 */
          const double w[4] = {cw[CL(-2)], cw[CL(-1)], cw[CL(0)], cw[CL(1)]};
          double c[4];
          lagrange_coeff(w, 3, 1, 1, c);
          for (int e = 0; e < neq; ++e)
            state[e] = c[0] * b->state[neq * CL(-2) + e] + c[1] * b->state[neq * CL(-1) + e] +
                       c[2] * b->state[neq * CL(0) + e] + c[3] * b->state[neq * CL(1) + e];
          mu = c[0] * b->viscosity[CL(-2)] + c[1] * b->viscosity[CL(-1)] +
               c[2] * b->viscosity[CL(0)] + c[3] * b->viscosity[CL(1)];
          const double w2[2] = {w[1], w[2]};
          double c2[2];
          lagrange_coeff(w2, 1, 0, 0, c2);
          /* turbulence variables stay second order (reconstruction.hpp:366-376) */
          for (int e = ns + 4; e < neq; ++e)
            state[e] = c2[0] * b->state[neq * CL(0) + e] + c2[1] * b->state[neq * CL(-1) + e];
          if (b->wallDist) wDist = c2[0] * b->wallDist[CL(0)] + c2[1] * b->wallDist[CL(-1)];
        }
        limit_turb(h, state);
        if (wDist < 0.0 && wDist > -1.0e-10) wDist = 0.0; /* WALL_DIST_NEG_TOL, :1349-1351 */
        if (rans) eddy_visc_and_blending(h, state, vg, kg, wg, mu, wDist, &mut, &f1, &f2);
        /* viscousFlux::CalcFlux / CalcWallFlux; ref: src/viscousFlux.cpp:58-135,137-196 */
        const double mus = h->cfg.nondimScaling * mu;
        const double muts = h->cfg.nondimScaling * mut;
        double tau[3];
        tau_normal(vg, fa, mus, muts, tau);
        flux[ns] = tau[0];
        flux[ns + 1] = tau[1];
        flux[ns + 2] = tau[2];
        const double t = temperature_of(h, state);
        const double kcond = eff_conductivity(h, t, state);
        double kt = 0.0;
        if (rans) { /* sutherland::TurbConductivity: mut cp / Prt, include/transport.hpp:132-137 */
          double mf[AITHER_MAX_SPECIES];
          mass_fractions(h, state, mf);
          kt = muts * cp_mix(h, mf) / turb_prandtl(h);
        }
        /* species diffusion with the zero-net-mass-flux rescale; not at low-Re wall faces,
         * whose CalcWallFlux has no species terms (ref: src/viscousFlux.cpp:84-105,137-196) */
        double speciesEnthalpyTerm = 0.0;
        if (ns > 1 && !isLowReWall) {
          /* schmidt::DiffCoeff (include/diffusion.hpp:100-105; Sc_t = 0.7, turbulence.hpp:71);
           * diffusionModel none: 0 */
          const double dc =
              h->cfg.schmidt > 0.0 ? mus / h->cfg.schmidt + muts / 0.7 : 0.0;
          double posDiff = 0.0, negDiff = 0.0;
          for (int ss = 0; ss < ns; ++ss) {
            flux[ss] = dc * dot3(mg[ss], fa);
            negDiff -= flux[ss] < 0.0 ? flux[ss] : 0.0;
            posDiff += flux[ss] > 0.0 ? flux[ss] : 0.0;
          }
          const double posDiffFac = posDiff > negDiff ? negDiff / posDiff : 1.0;
          const double negDiffFac = negDiff > posDiff ? posDiff / negDiff : 1.0;
          const double tf = temperature_of(h, state);
          const double vel = vel_mag(h, state);
          for (int ss = 0; ss < ns; ++ss) {
            flux[ss] *= flux[ss] > 0.0 ? posDiffFac : negDiffFac;
            const double hs = (h->cfg.hf[ss] +
                               h->cfg.gasConstant[ss] * (h->cfg.n[ss] + 1.0) * tf) +
                              0.5 * vel * vel;
            speciesEnthalpyTerm += flux[ss] * hs;
          }
        }
        flux[ns + 3] = (tau[0] * state[ns] + tau[1] * state[ns + 1] + tau[2] * state[ns + 2]) +
                       (kcond + kt) * (tg[0] * fa[0] + tg[1] * fa[1] + tg[2] * fa[2]) +
                       speciesEnthalpyTerm;
        if (rans) {
          const double mutt = is_sst(h) ? muts
                                        : h->cfg.nondimScaling * eddy_visc_no_lim(h, state);
          flux[ns + 4] = (mus + sigma_k(h, f1) * mutt) * dot3(kg, fa);
          flux[ns + 5] = (mus + sigma_w(h, f1) * mutt) * dot3(wg, fa);
        }
        } /* not a wall-law face */
        /* residual: opposite sign to the inviscid flux; ref: :1392-1429 */
        if (fi > 0) {
          double *r = b->residual + neq * pidx(b, ii - di, jj - dj, kk - dk);
          for (int e = 0; e < neq; ++e) r[e] -= flux[e] * fa[3];
          if (!rans)
            for (int q = 0; q < 9; ++q) b->velGrad[9 * CL(-1) + q] += sixth * vg[q];
          for (int q = 0; q < 3; ++q)
            b->pressGrad[3 * pidx(b, ii - di, jj - dj, kk - dk) + q] += sixth * pg[q];
          if (h->cfg.isBlockMatrix) { /* ref: src/procBlock.cpp:1420-1428 */
            double J[MAXJAC];
            approx_tsl_jacobian(h, state, mu, mut, f1, fa, proj_c2c_dist(b, d, ii, jj, kk), 1,
                                vg, J);
            double *a = b->a + h->asz * pidx(b, ii - di, jj - dj, kk - dk);
            for (int q = 0; q < h->asz; ++q) a[q] -= J[q];
          }
          if (rans) {
            const long c = CL(-1), pp = pidx(b, ii - di, jj - dj, kk - dk);
            for (int q = 0; q < 9; ++q) b->velGrad[9 * c + q] += sixth * vg[q];
            b->eddyVisc[c] += sixth * mut;
            for (int q = 0; q < 3; ++q) {
              b->tkeGrad[3 * pp + q] += sixth * kg[q];
              b->omegaGrad[3 * pp + q] += sixth * wg[q];
            }
            b->f1[c] += sixth * f1;
            b->f2[c] += sixth * f2;
          }
        }
        if (fi < nd[d]) {
          double *r = b->residual + neq * pidx(b, ii, jj, kk);
          for (int e = 0; e < neq; ++e) r[e] += flux[e] * fa[3];
          if (!rans)
            for (int q = 0; q < 9; ++q) b->velGrad[9 * CL(0) + q] += sixth * vg[q];
          for (int q = 0; q < 3; ++q) b->pressGrad[3 * pidx(b, ii, jj, kk) + q] += sixth * pg[q];
          if (rans) {
            const long c = CL(0), pp = pidx(b, ii, jj, kk);
            for (int q = 0; q < 9; ++q) b->velGrad[9 * c + q] += sixth * vg[q];
            b->eddyVisc[c] += sixth * mut;
            for (int q = 0; q < 3; ++q) {
              b->tkeGrad[3 * pp + q] += sixth * kg[q];
              b->omegaGrad[3 * pp + q] += sixth * wg[q];
            }
            b->f1[c] += sixth * f1;
            b->f2[c] += sixth * f2;
          }
          /* ViscCellSpectralRadius; ref: include/spectralRadius.hpp:94-124 */
          const double *s = b->state + neq * CL(0);
          const double *fu = farea(b, d, ii + di, jj + dj, kk + dk);
          const double fMag = 0.5 * (fa[3] + fu[3]);
          const double rho = rho_of(h, s);
          const double gam = gamma_of(h, s);
          const double a43 = 4.0 / (3.0 * rho), gr = gam / rho;
          const double maxTerm = a43 > gr ? a43 : gr;
          const double viscTerm =
              h->cfg.nondimScaling * (b->viscosity[CL(0)] / prandtl_of(gam) + mut / turb_prandtl(h));
          const double vsr = maxTerm * viscTerm * fMag * fMag / b->vol[CL(0)];
          /* turbModel::ViscCellSpecRad; ref: src/turbulence.cpp:500-511,783-794 */
          const double tvsr =
              rans ? turb_visc_spec_rad(h, s, fMag * fMag / b->vol[CL(0)], b->viscosity[CL(0)],
                                        mut, f1)
                   : 0.0;
          double *sp = b->specRad + 2 * pidx(b, ii, jj, kk);
          sp[0] += vsr * viscCoeff;
          sp[1] += tvsr * viscCoeff;
          if (!h->cfg.isBlockMatrix) {
            double *a = b->a + h->asz * pidx(b, ii, jj, kk);
            a[0] += 2.0 * vsr;
            if (rans) a[1] += 2.0 * tvsr;
          } else { /* ref: src/procBlock.cpp:1481-1489 */
            double J[MAXJAC];
            approx_tsl_jacobian(h, state, mu, mut, f1, fa, proj_c2c_dist(b, d, ii, jj, kk), 0,
                                vg, J);
            double *a = b->a + h->asz * pidx(b, ii, jj, kk);
            for (int q = 0; q < h->asz; ++q) a[q] += J[q];
          }
        }
#undef CL
      }
}

/* procBlock::CalcSrcTerms (RANS part); ref: src/procBlock.cpp:5956-6025,
 * src/source.cpp:64-82 */
static void calc_src_terms(orc_level *h, orc_block *b) {
  const int neq = h->neq, ns = h->ns;
  for (int kk = 0; kk < b->nk; ++kk)
    for (int jj = 0; jj < b->nj; ++jj)
      for (int ii = 0; ii < b->ni; ++ii) {
        const long c = cidx(b, ii, jj, kk), pp = pidx(b, ii, jj, kk);
        const double *s = b->state + neq * c;
        double src[2], beta = 0.0;
        calc_turb_src(h, s, b->velGrad + 9 * c, b->tkeGrad + 3 * pp, b->omegaGrad + 3 * pp,
                      b->eddyVisc[c], b->f1[c], src, &beta);
        const double turbSpecRad = turb_src_spec_rad(h, s, b->vol[c]);
        b->specRad[2 * pp + 1] -= turbSpecRad;
        if (!h->cfg.isBlockMatrix) {
          b->a[h->asz * pp + 1] -= turbSpecRad;
        } else { /* TurbSrcJac; ref: src/turbulence.cpp:445-458,706-720 */
          const int fs = ns + 4;
          const double invScaling = 1.0 / h->cfg.nondimScaling;
          double *a = b->a + h->asz * pp + fs * fs;
          const double j00 = is_sst(h)
                                 ? -2.0 * SST_BETASTAR * omega_of(h, s) * 1.0 * b->vol[c] * invScaling
                                 : -2.0 * KW_BETASTAR * omega_of(h, s) * b->vol[c] * invScaling;
          const double j11 = -2.0 * beta * omega_of(h, s) * b->vol[c] * invScaling;
          a[0] -= j00;
          a[1] -= 0.0;
          a[2] -= 0.0;
          a[3] -= j11;
        }
        double *r = b->residual + neq * pp;
        for (int e = 0; e < neq; ++e) {
          const double sv = e >= ns + 4 ? src[e - ns - 4] : 0.0;
          r[e] -= sv * b->vol[c];
        }
      }
}

/* ------------------------------------------------------------------------ */
/* connection (interblock / periodic) ghost swaps                            */
typedef struct {
  int boundary[2], d1s[2], d1e[2], d2s[2], d2e[2], cs[2], border[8], orient;
} orc_conn;
static void conn_load(const aither_conn *c, orc_conn *o) {
  for (int s = 0; s < 2; ++s) {
    o->boundary[s] = c->boundary[s];
    o->d1s[s] = c->d1Start[s];
    o->d1e[s] = c->d1End[s];
    o->d2s[s] = c->d2Start[s];
    o->d2e[s] = c->d2End[s];
    o->cs[s] = c->constSurf[s];
  }
  for (int q = 0; q < 8; ++q) o->border[q] = c->patchBorder[q];
  o->orient = c->orientation;
}
#define ISWAP(a, b) do { int t_ = (a); (a) = (b); (b) = t_; } while (0)
/* ref: src/boundaryConditions.cpp:341-364 */
static void conn_swap_order(orc_conn *o) {
  ISWAP(o->boundary[0], o->boundary[1]);
  ISWAP(o->d1s[0], o->d1s[1]);
  ISWAP(o->d1e[0], o->d1e[1]);
  ISWAP(o->d2s[0], o->d2s[1]);
  ISWAP(o->d2e[0], o->d2e[1]);
  ISWAP(o->cs[0], o->cs[1]);
  for (int q = 0; q < 4; ++q) ISWAP(o->border[q], o->border[q + 4]);
  if (o->orient == 4) o->orient = 5;
  else if (o->orient == 5) o->orient = 4;
}
/* ref: src/boundaryConditions.cpp:833-858 */
static void conn_adjust_for_slice(orc_conn *o, int blkFirst, int numG) {
  if (!blkFirst) conn_swap_order(o);
  const int blkStart = (o->boundary[0] % 2 == 0) ? o->cs[0] : -numG;
  o->cs[1] = 0;
  o->cs[0] = blkStart;
  o->d1e[1] = o->d1e[1] - o->d1s[1] + 2 * numG;
  o->d1e[0] = o->d1e[0] + numG;
  o->d1s[1] = 0;
  o->d1s[0] = o->d1s[0] - numG;
  o->d2e[1] = o->d2e[1] - o->d2s[1] + 2 * numG;
  o->d2e[0] = o->d2e[0] + numG;
  o->d2s[1] = 0;
  o->d2s[0] = o->d2s[0] - numG;
}
/* ref: src/boundaryConditions.cpp:1016-1150; side s, lo/hi in (i, j, k) */
static void conn_slice_indices(const orc_conn *o, int s, int numG, int *lo,
                               int *hi) {
  const int upLowFac = (o->boundary[s] % 2 == 0) ? -numG : 0;
  int d3, d1, d2;
  if (o->boundary[s] <= 2) { d3 = 0; d1 = 1; d2 = 2; }
  else if (o->boundary[s] <= 4) { d3 = 1; d1 = 2; d2 = 0; }
  else { d3 = 2; d1 = 0; d2 = 1; }
  lo[d3] = o->cs[s] + upLowFac;
  hi[d3] = lo[d3] + numG;
  lo[d1] = o->d1s[s] - numG;
  hi[d1] = o->d1e[s] + numG;
  lo[d2] = o->d2s[s] - numG;
  hi[d2] = o->d2e[s] + numG;
}
/* ref: src/boundaryConditions.cpp:3006-3181 (GetSwapLoc) */
static void get_swap_loc(int l1, int l2, int l3, int numGhosts,
                         const orc_conn *o, int d3, int first, int *loc) {
  const int ot = o->orient;
  if (first) {
    const int lower = o->cs[0] == 0; /* IsLowerFirst */
    const int n3 = lower ? l3 - numGhosts : o->cs[0] + l3;
    if (o->boundary[0] <= 2) {
      loc[1] = o->d1s[0] + l1; loc[2] = o->d2s[0] + l2; loc[0] = n3;
    } else if (o->boundary[0] <= 4) {
      loc[2] = o->d1s[0] + l1; loc[0] = o->d2s[0] + l2; loc[1] = n3;
    } else {
      loc[0] = o->d1s[0] + l1; loc[1] = o->d2s[0] + l2; loc[2] = n3;
    }
    return;
  }
  const int lowerSecond = o->cs[1] == 0; /* IsLowerSecond */
  const int llOrUu = (o->boundary[0] + o->boundary[1]) % 2 == 0;
  int n3;
  if (llOrUu)
    n3 = lowerSecond ? d3 - l3 - 1 : o->cs[1] + d3 - l3 - 1;
  else
    n3 = lowerSecond ? l3 - numGhosts : o->cs[1] + l3;
  const int swapped = ot == 2 || ot == 4 || ot == 5 || ot == 7;
  if (o->boundary[1] <= 2) { /* i-patch: dir 1 = j, dir 2 = k */
    if (swapped) {
      loc[2] = (ot == 5 || ot == 7) ? o->d2e[1] - 1 - l1 : o->d2s[1] + l1;
      loc[1] = (ot == 4 || ot == 7) ? o->d1e[1] - 1 - l2 : o->d1s[1] + l2;
    } else {
      loc[1] = (ot == 6 || ot == 8) ? o->d1e[1] - 1 - l1 : o->d1s[1] + l1;
      loc[2] = (ot == 3 || ot == 8) ? o->d2e[1] - 1 - l2 : o->d2s[1] + l2;
    }
    loc[0] = n3;
  } else if (o->boundary[1] <= 4) { /* j-patch: dir 1 = k, dir 2 = i */
    if (swapped) {
      loc[0] = (ot == 5 || ot == 7) ? o->d2e[1] - 1 - l1 : o->d2s[1] + l1;
      loc[2] = (ot == 4 || ot == 7) ? o->d1e[1] - 1 - l2 : o->d1s[1] + l2;
    } else {
      loc[2] = (ot == 3 || ot == 8) ? o->d1e[1] - 1 - l1 : o->d1s[1] + l1;
      loc[0] = (ot == 6 || ot == 8) ? o->d2e[1] - 1 - l2 : o->d2s[1] + l2;
    }
    loc[1] = n3;
  } else { /* k-patch: dir 1 = i, dir 2 = j */
    if (swapped) {
      loc[1] = (ot == 5 || ot == 7) ? o->d2e[1] - 1 - l1 : o->d2s[1] + l1;
      loc[0] = (ot == 4 || ot == 7) ? o->d1e[1] - 1 - l2 : o->d1s[1] + l2;
    } else {
      loc[0] = (ot == 3 || ot == 8) ? o->d1e[1] - 1 - l1 : o->d1s[1] + l1;
      loc[1] = (ot == 6 || ot == 8) ? o->d2e[1] - 1 - l2 : o->d2s[1] + l2;
    }
    loc[2] = n3;
  }
}
/* copy of a box of a padded array (multiArray3d::Slice) */
static double *take_slice(const orc_block *b, const double *arr, int nc,
                          const int *lo, const int *hi) {
  const int n0 = hi[0] - lo[0], n1 = hi[1] - lo[1], n2 = hi[2] - lo[2];
  double *s = (double *)malloc(sizeof(double) * (size_t)n0 * n1 * n2 * nc);
  for (int k = 0; k < n2; ++k)
    for (int j = 0; j < n1; ++j)
      for (int i = 0; i < n0; ++i)
        memcpy(s + (size_t)nc * (i + (size_t)n0 * (j + (size_t)n1 * k)),
               arr + nc * cidx(b, lo[0] + i, lo[1] + j, lo[2] + k),
               sizeof(double) * nc);
  return s;
}
/* ref: include/multiArray3d.hpp:876-926 (InsertSlice); `o` already adjusted */
static void insert_slice(const orc_block *b, double *arr, int nc,
                         const double *slice, const int *sn,
                         const orc_conn *o, int d3) {
  const int adjS1 = o->border[0] ? b->g : 0, adjE1 = o->border[1] ? b->g : 0;
  const int adjS2 = o->border[2] ? b->g : 0, adjE2 = o->border[3] ? b->g : 0;
  const int len1 = o->d1e[0] - o->d1s[0], len2 = o->d2e[0] - o->d2s[0];
  for (int l3 = 0; l3 < d3; ++l3)
    for (int l2 = adjS2; l2 < len2 - adjE2; ++l2)
      for (int l1 = adjS1; l1 < len1 - adjE1; ++l1) {
        int a[3], s[3];
        get_swap_loc(l1, l2, l3, b->g, o, d3, 1, a);
        get_swap_loc(l1, l2, l3, 0, o, d3, 0, s);
        memcpy(arr + nc * cidx(b, a[0], a[1], a[2]),
               slice + (size_t)nc * (s[0] + (size_t)sn[0] *
                                                (s[1] + (size_t)sn[1] * s[2])),
               sizeof(double) * nc);
      }
}
/* ref: include/multiArray3d.hpp:790-823 (SwapSliceLocal), connection by
 * connection in list order (src/gridLevel.cpp:297-312, src/utility.cpp:400-423);
 * which = 0: state, 1: implicit update x */
static void swap_connections(orc_level *h, int which) {
  /* 2..5: eddy viscosity, f1, f2, velocity gradient
   * (ref: src/procBlock.cpp:3064-3085, src/gridLevel.cpp:321-370) */
  const int nc = which <= 1 ? h->neq : (which == 5 ? 9 : 1);
  for (int c = 0; c < h->nconn; ++c) {
    orc_conn o;
    conn_load(&h->conn[c], &o);
    orc_block *b1 = &h->blk[h->conn[c].localBlock[0]];
    orc_block *b2 = &h->blk[h->conn[c].localBlock[1]];
#define SWAP_ARR(bp)                                                       \
  (which == 0 ? (bp)->state                                                \
              : (which == 1 ? (bp)->x                                      \
                            : (which == 2 ? (bp)->eddyVisc                 \
                                          : (which == 3 ? (bp)->f1         \
                                                        : (which == 4 ? (bp)->f2 : (bp)->velGrad)))))
    double *a1 = SWAP_ARR(b1);
    double *a2 = SWAP_ARR(b2);
#undef SWAP_ARR
    int lo1[3], hi1[3], lo2[3], hi2[3], sn1[3], sn2[3];
    conn_slice_indices(&o, 0, b1->g, lo1, hi1);
    conn_slice_indices(&o, 1, b2->g, lo2, hi2);
    for (int q = 0; q < 3; ++q) { sn1[q] = hi1[q] - lo1[q]; sn2[q] = hi2[q] - lo2[q]; }
    double *s1 = take_slice(b1, a1, nc, lo1, hi1);
    double *s2 = take_slice(b2, a2, nc, lo2, hi2);
    orc_conn c1 = o, c2 = o;
    conn_adjust_for_slice(&c1, 0, b1->g);
    conn_adjust_for_slice(&c2, 1, b2->g);
    insert_slice(b1, a1, nc, s2, sn2, &c2, b2->g);
    insert_slice(b2, a2, nc, s1, sn1, &c1, b1->g);
    free(s1);
    free(s2);
  }
}

/* ------------------------------------------------------------------------ */
void orc_get_boundary_conditions(orc_level *h) {
  /* ref: src/gridLevel.cpp:287-319 */
  for (int bb = 0; bb < h->nblk; ++bb) assign_inviscid_ghosts(h, &h->blk[bb]);
  swap_connections(h, 0);
  for (int bb = 0; bb < h->nblk; ++bb) assign_ghost_edges(h, &h->blk[bb], 0);
}

void orc_calc_residual(orc_level *h) {
  /* ref: src/procBlock.cpp:6111-6147, src/gridLevel.cpp:372-400 */
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    const long nc = (long)b->ni * b->nj * b->nk;
    const long np = (long)b->NI * b->NJ * b->NK;
    memset(b->residual, 0, sizeof(double) * nc * h->neq);
    memset(b->specRad, 0, sizeof(double) * nc * 2);
    if (h->cfg.isViscous) memset(b->velGrad, 0, sizeof(double) * np * 9);
    memset(b->pressGrad, 0, sizeof(double) * nc * 3);
    if (is_rans(h)) { /* ResetGradients, ResetTurbVars; ref: src/procBlock.cpp:961-981 */
      memset(b->tkeGrad, 0, sizeof(double) * nc * 3);
      memset(b->omegaGrad, 0, sizeof(double) * nc * 3);
      memset(b->eddyVisc, 0, sizeof(double) * np);
      memset(b->f1, 0, sizeof(double) * np);
      memset(b->f2, 0, sizeof(double) * np);
    }
    calc_inv_flux(h, b, 0);
    calc_inv_flux(h, b, 1);
    calc_inv_flux(h, b, 2);
    if (h->cfg.isViscous) { /* ref: src/procBlock.cpp:6125-6137 */
      assign_viscous_ghosts(h, b);
      update_aux(h, b);
      calc_visc_flux(h, b, 0);
      calc_visc_flux(h, b, 1);
      calc_visc_flux(h, b, 2);
    } else {
      update_aux(h, b);
      /* gradient-only pass of Euler runs (CalcGradsI/J/K; ref: src/procBlock.cpp:6138-6146,
       * :5790-5945): the cell averages are read by the non-reflecting BCs only, so the pass
       * is skipped when there is none */
      int nonRefl = 0;
      for (int q = 0; q < h->cfg.numBCStates; ++q) nonRefl |= h->cfg.bcStates[q].isNonreflecting;
      if (nonRefl) {
        memset(b->velGrad, 0, sizeof(double) * np * 9);
        const int nd[3] = {b->ni, b->nj, b->nk};
        const double sixth = 1.0 / 6.0;
        for (int d = 0; d < 3; ++d) {
          const int di = d == 0, dj = d == 1, dk = d == 2;
          for (int kk = 0; kk < b->nk + dk; ++kk)
            for (int jj = 0; jj < b->nj + dj; ++jj)
              for (int ii = 0; ii < b->ni + di; ++ii) {
                const int fi = d == 0 ? ii : (d == 1 ? jj : kk);
                double vg[9], tg[3], kg[3], wg[3], mg[AITHER_MAX_SPECIES][3];
                face_gradients(h, b, d, ii, jj, kk, vg, tg, kg, wg, mg);
                if (fi > 0) {
                  const long c = cidx(b, ii - di, jj - dj, kk - dk);
                  const long pp = pidx(b, ii - di, jj - dj, kk - dk);
                  for (int q = 0; q < 9; ++q) b->velGrad[9 * c + q] += sixth * vg[q];
                  for (int q = 0; q < 3; ++q) b->pressGrad[3 * pp + q] += sixth * g_face_press_grad[q];
                }
                if (fi < nd[d]) {
                  const long c = cidx(b, ii, jj, kk), pp = pidx(b, ii, jj, kk);
                  for (int q = 0; q < 9; ++q) b->velGrad[9 * c + q] += sixth * vg[q];
                  for (int q = 0; q < 3; ++q) b->pressGrad[3 * pp + q] += sixth * g_face_press_grad[q];
                }
              }
        }
      }
    }
  }
  if (h->cfg.isViscous && !is_rans(h)) swap_connections(h, 5);
  if (is_rans(h)) {
    /* SwapEddyViscAndGradients, SwapTurbVars, then the source terms;
     * ref: src/gridLevel.cpp:386-399 */
    swap_connections(h, 5);
    swap_connections(h, 2);
    swap_connections(h, 3);
    swap_connections(h, 4);
    for (int bb = 0; bb < h->nblk; ++bb) calc_src_terms(h, &h->blk[bb]);
  }
}

void orc_calc_time_step(orc_level *h, double cfl) {
  /* ref: src/procBlock.cpp:782-821 */
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii) {
          const long p = pidx(b, ii, jj, kk);
          if (h->cfg.dtNondim > 0.0) {
            b->dt[p] = h->cfg.dtNondim;
          } else {
            const double *sp = b->specRad + 2 * p;
            const double mx = sp[0] > sp[1] ? sp[0] : sp[1];
            b->dt[p] = cfl * (b->vol[cidx(b, ii, jj, kk)] / mx);
          }
        }
  }
}

/* ref: src/procBlock.cpp:1010-1012 */
static double sol_delta_n_coeff(const orc_level *h, const orc_block *b, int ii,
                                int jj, int kk) {
  return (b->vol[cidx(b, ii, jj, kk)] * (1.0 + h->cfg.zeta)) /
         (b->dt[pidx(b, ii, jj, kk)] * h->cfg.theta);
}

static void matrix_inverse(double *mat, int size);
void orc_invert_diagonal(orc_level *h) {
  /* ref: src/linearSolver.cpp:146-188 (scalar diagonal) */
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii) {
          const long p = pidx(b, ii, jj, kk);
          double diagVolTime = sol_delta_n_coeff(h, b, ii, jj, kk);
          if (h->cfg.dualTimeCFL > 0.0) {
            const double *sp = b->specRad + 2 * p;
            diagVolTime += (sp[0] > sp[1] ? sp[0] : sp[1]) / h->cfg.dualTimeCFL;
          }
          if (h->cfg.isBlockMatrix) {
            /* MultiplyOnDiagonal / AddOnDiagonal / Inverse of the flow and turbulence
             * blocks; ref: include/matMultiArray3d.hpp:109-122 */
            const int fs = h->ns + 4, nt = h->nt;
            double *a = b->a + h->asz * p, *ai = b->ainv + h->asz * p;
            for (int r = 0; r < fs; ++r) a[r * fs + r] *= h->cfg.matrixRelaxation;
            for (int r = 0; r < nt; ++r) a[fs * fs + r * nt + r] *= h->cfg.matrixRelaxation;
            for (int r = 0; r < fs; ++r) a[r * fs + r] += diagVolTime;
            for (int r = 0; r < nt; ++r) a[fs * fs + r * nt + r] += diagVolTime;
            memcpy(ai, a, sizeof(double) * h->asz);
            matrix_inverse(ai, fs);
            if (nt > 0) matrix_inverse(ai + fs * fs, nt);
            continue;
          }
          for (int q = 0; q < h->asz; ++q) {
            b->a[h->asz * p + q] *= h->cfg.matrixRelaxation;
            b->a[h->asz * p + q] += diagVolTime;
            b->ainv[h->asz * p + q] = 1.0 * (1.0 / b->a[h->asz * p + q]);
          }
        }
  }
}

/* b = -R/theta + SolDeltaNm1 - SolDeltaMmN; ref: src/procBlock.cpp:1014-1034,
 * src/linearSolver.cpp:124-129 */
static void rhs_b(const orc_level *h, const orc_block *b, int ii, int jj,
                  int kk, double *out) {
  const int neq = h->neq;
  const long p = pidx(b, ii, jj, kk);
  const double thetaInv = 1.0 / h->cfg.theta;
  double cons[MAXEQ];
  prim_to_cons(h, b->state + neq * cidx(b, ii, jj, kk), cons);
  const double coeff = sol_delta_n_coeff(h, b, ii, jj, kk);
  for (int e = 0; e < neq; ++e) {
    double nm1 = 0.0;
    if (h->cfg.isMultilevelTime) {
      const double c1 = (b->vol[cidx(b, ii, jj, kk)] * h->cfg.zeta) /
                        (b->dt[p] * h->cfg.theta);
      nm1 = c1 * (b->consN[neq * p + e] - b->consNm1[neq * p + e]);
    }
    const double mmn = coeff * (cons[e] - b->consN[neq * p + e]);
    out[e] = -thetaInv * b->residual[neq * p + e] + nm1 - mmn;
  }
}

static void block_mult(const orc_level *h, const double *a, const double *v, double *out);
static void diag_mult(const orc_level *h, const double *a, const double *v,
                      double *out) {
  if (h->cfg.isBlockMatrix) {
    double tmp[MAXEQ];
    block_mult(h, a, v, tmp);
    for (int e = 0; e < h->neq; ++e) out[e] = tmp[e];
    return;
  }
  /* scalar: ref include/fluxJacobian.hpp:67-73 */
  for (int e = 0; e < h->ns + 4; ++e) out[e] = v[e] * a[0];
  for (int e = h->ns + 4; e < h->neq; ++e) out[e] = v[e] * a[1];
}

void orc_initialize_matrix_update(orc_level *h) {
  /* ref: src/linearSolver.cpp:111-144 */
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    const long np = (long)b->NI * b->NJ * b->NK * h->neq;
    if (!h->cfg.matrixRequiresInit) {
      memset(b->x, 0, sizeof(double) * np);
      continue;
    }
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii) {
          double rb[MAXEQ];
          rhs_b(h, b, ii, jj, kk, rb);
          diag_mult(h, b->ainv + h->asz * pidx(b, ii, jj, kk), rb,
                    b->x + h->neq * cidx(b, ii, jj, kk));
        }
  }
}

/* ------------------------------------------------------------------------ */
/* block flux Jacobians (blusgs / bdplur): flow block fs x fs row-major, then  */
/* the turbulence block nt x nt (ref: include/fluxJacobian.hpp:62-75)          */
static double energy_of(const orc_level *h, const double *s);
/* fluxJacobian::InvFluxJacobian; ref: include/fluxJacobian.hpp:486-562 */
static void inv_flux_jacobian(const orc_level *h, const double *s,
                              const double *area, double *J) {
  const int ns = h->ns, fs = ns + 4, nt = h->nt;
  for (int q = 0; q < fs * fs + nt * nt; ++q) J[q] = 0.0;
  const double *n = area;
  const double u = s[ns], v = s[ns + 1], w = s[ns + 2];
  const double velNorm = u * n[0] + v * n[1] + w * n[2];
  double mf[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  const double gamma = gamma_of(h, s);
  const double gm1 = gamma - 1.0;
  const double phi = 0.5 * gm1 * (u * u + v * v + w * w);
  const double a1 = gamma * energy_of(h, s) - phi;
  const double a3 = gamma - 2.0;
#define FJ(r, c) J[(r)*fs + (c)]
  for (int ii = 0; ii < ns; ++ii) {
    for (int jj = 0; jj < ns; ++jj)
      FJ(ii, jj) = velNorm * ((ii == jj ? 1.0 : 0.0) - mf[ii]);
    FJ(ii, ns + 0) = mf[ii] * n[0];
    FJ(ii, ns + 1) = mf[ii] * n[1];
    FJ(ii, ns + 2) = mf[ii] * n[2];
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
  for (int q = 0; q < fs * fs; ++q) J[q] *= 0.5 * area[3];
  if (nt > 0) { /* 0.5 * InviscidConvJacobian; ref: src/turbulence.cpp:126-136 */
    const double diag = velNorm * area[3];
    J[fs * fs + 0] = 0.5 * diag;
    J[fs * fs + 3] = 0.5 * diag;
  }
}
/* fluxJacobian::RusanovFluxJacobian; ref: include/fluxJacobian.hpp:446-483 */
static void rusanov_flux_jacobian(const orc_level *h, const double *s,
                                  const double *area, int positive, double *J) {
  const int fs = h->ns + 4, nt = h->nt;
  const double specRad = inv_face_spec_rad(h, s, area);
  inv_flux_jacobian(h, s, area, J);
  double D[MAXJAC];
  for (int q = 0; q < fs * fs + nt * nt; ++q) D[q] = 0.0;
  for (int r = 0; r < fs; ++r) D[r * fs + r] = 1.0 * specRad;
  if (nt > 0) { /* 0.5 * InviscidDissJacobian; ref: src/turbulence.cpp:138-148 */
    const double velNorm = dot3(s + h->ns, area);
    const double diag = fabs(velNorm) * area[3];
    D[fs * fs + 0] = 0.5 * diag;
    D[fs * fs + 3] = 0.5 * diag;
  }
  for (int q = 0; q < fs * fs + nt * nt; ++q) J[q] = positive ? J[q] + D[q] : J[q] - D[q];
}
/* fluxJacobian::ApproxTSLJacobian; ref: include/fluxJacobian.hpp:664-758
 * (DelprimitiveDelConservative :610-661, MatrixMultiply src/matrix.cpp:197-212) */
static void approx_tsl_jacobian(const orc_level *h, const double *s, double lamVisc,
                                double turbVisc, double f1, const double *area,
                                double dist, int left, const double vGrad[9],
                                double *J) {
  const int ns = h->ns, fs = ns + 4, nt = h->nt;
  double A[MAXFS * MAXFS], P[MAXFS * MAXFS];
  for (int q = 0; q < fs * fs; ++q) A[q] = P[q] = 0.0;
  for (int q = 0; q < fs * fs + nt * nt; ++q) J[q] = 0.0;
  const double t = temperature_of(h, s);
  const double mu = h->cfg.nondimScaling * lamVisc;
  const double mut = h->cfg.nondimScaling * turbVisc;
  const double *n = area;
  const double u = s[ns], v = s[ns + 1], w = s[ns + 2];
  const double velNorm = u * n[0] + v * n[1] + w * n[2];
  double mf[AITHER_MAX_SPECIES];
  mass_fractions(h, s, mf);
  const double rho = rho_of(h, s);
  const double k = eff_conductivity(h, t, s);
  const double kt = mut * cp_mix(h, mf) / turb_prandtl(h);
  double tauNorm[3];
  tau_normal(vGrad, n, mu, mut, tauNorm);
  const double fac = left ? -1.0 : 1.0;
  const double third = 1.0 / 3.0;
  if (ns != 1) {
    fprintf(stderr, "oracle: species diffusion jacobian is not restated (ns > 1)\n");
    abort();
  }
#define FA(r, c) A[(r)*fs + (c)]
  for (int ii = 0; ii < ns; ++ii) {
    for (int jj = 0; jj < ns; ++jj)
      FA(ii, jj) = 0.0 * ((ii == jj ? 1.0 : 0.0) - mf[ii]) / ((mu + mut) * rho);
    const double speciesEnthalpy = FA(ii, ii) * 0.0;
    FA(ns + 3, ii) = -(k + kt) * t / ((mu + mut) * rho) + speciesEnthalpy;
  }
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
  for (int q = 0; q < fs * fs; ++q) A[q] *= area[3] * (mu + mut) / dist;
  /* prim2Cons */
  const double gm1 = gamma_of(h, s) - 1.0;
  const double invRho = 1.0 / rho;
#define FP(r, c) P[(r)*fs + (c)]
  for (int ii = 0; ii < ns; ++ii) {
    FP(ii, ii) = 1.0;
    FP(ns + 0, ii) = -invRho * u;
    FP(ns + 1, ii) = -invRho * v;
    FP(ns + 2, ii) = -invRho * w;
    FP(ns + 3, ii) = 0.5 * gm1 * (u * u + v * v + w * w);
  }
  FP(ns, ns) = invRho;
  FP(ns + 3, ns) = -gm1 * u;
  FP(ns + 1, ns + 1) = invRho;
  FP(ns + 3, ns + 1) = -gm1 * v;
  FP(ns + 2, ns + 2) = invRho;
  FP(ns + 3, ns + 2) = -gm1 * w;
  FP(ns + 3, ns + 3) = gm1;
#undef FP
  for (int cc = 0; cc < fs; ++cc)
    for (int rr = 0; rr < fs; ++rr)
      for (int ii = 0; ii < fs; ++ii) J[rr * fs + ii] += A[rr * fs + cc] * P[cc * fs + ii];
  if (nt > 0) { /* fac * turbModel::ViscJac; ref: src/turbulence.cpp:485-498,768-781 */
    const double length = area[3] / dist;
    const double mt = is_sst(h) ? turbVisc : eddy_visc_no_lim(h, s);
    J[fs * fs + 0] = fac * (h->cfg.nondimScaling * length / rho * (lamVisc + sigma_k(h, f1) * mt));
    J[fs * fs + 3] = fac * (h->cfg.nondimScaling * length / rho * (lamVisc + sigma_w(h, f1) * mt));
  }
}
/* squareMatrix inverse by Gauss-Jordan elimination with partial pivoting, the
 * reference's operation order; ref: src/matrix.cpp:57-107 */
static void matrix_inverse(double *mat, int size) {
  double I[MAXFS * MAXFS];
  for (int r = 0; r < size; ++r)
    for (int c = 0; c < size; ++c) I[r * size + c] = r == c ? 1.0 : 0.0;
  for (int cPivot = 0, r = 0; r < size; ++r, ++cPivot) {
    double maxVal = 0.0;
    int rPivot = 0;
    for (int ii = r; ii <= size - 1; ++ii)
      if (fabs(mat[ii * size + cPivot]) > maxVal) {
        maxVal = fabs(mat[ii * size + cPivot]);
        rPivot = ii;
      }
    if (r != rPivot)
      for (int c = 0; c < size; ++c) {
        double tmp = mat[r * size + c];
        mat[r * size + c] = mat[rPivot * size + c];
        mat[rPivot * size + c] = tmp;
        tmp = I[r * size + c];
        I[r * size + c] = I[rPivot * size + c];
        I[rPivot * size + c] = tmp;
      }
    if (r != 0)
      for (int ii = 0; ii < cPivot; ++ii) {
        const double factor = mat[r * size + ii] / mat[ii * size + ii];
        for (int c = 0; c < size; ++c) {
          mat[r * size + c] = mat[r * size + c] - factor * mat[ii * size + c];
          I[r * size + c] = I[r * size + c] - factor * I[ii * size + c];
        }
      }
    if (mat[r * size + cPivot] == 0.0) {
      fprintf(stderr, "oracle: singular matrix in Gauss-Jordan elimination\n");
      abort();
    }
    const double normFactor = 1.0 / mat[r * size + cPivot];
    for (int c = cPivot; c < size; ++c) mat[r * size + c] *= normFactor;
    for (int c = 0; c < size; ++c) I[r * size + c] *= normFactor;
  }
  for (int cPivot = size - 2, r = size - 2; r >= 0; --r, --cPivot)
    for (int ii = size - 1; ii > cPivot; --ii) {
      const double factor = mat[r * size + ii];
      for (int c = 0; c < size; ++c) {
        mat[r * size + c] = mat[r * size + c] - factor * mat[ii * size + c];
        I[r * size + c] = I[r * size + c] - factor * I[ii * size + c];
      }
    }
  for (int q = 0; q < size * size; ++q) mat[q] = I[q];
}
/* block ArrayMult; ref: include/fluxJacobian.hpp:76-88 */
static void block_mult(const orc_level *h, const double *a, const double *v, double *out) {
  const int fs = h->ns + 4, nt = h->nt;
  for (int e = 0; e < h->neq; ++e) out[e] = 0.0;
  for (int rr = 0; rr < fs; ++rr)
    for (int cc = 0; cc < fs; ++cc) out[rr] += a[rr * fs + cc] * v[cc];
  for (int rr = 0; rr < nt; ++rr)
    for (int cc = 0; cc < nt; ++cc) out[fs + rr] += a[fs * fs + rr * nt + cc] * v[fs + cc];
}

/* ref: src/fluxJacobian.cpp:122-162 (RusanovScalarOffDiagonal) */
static void offdiag_scalar(const orc_level *h, const double *state,
                           const double *du, const double *fArea, int positive,
                           double mu, double mut, double f1, double dist,
                           double *out) {
  const int neq = h->neq;
  double su[MAXEQ], fo[MAXEQ], fn[MAXEQ];
  update_prim_with_cons(h, state, du, su);
  physical_flux(h, state, fArea, fo);
  physical_flux(h, su, fArea, fn);
  double sr = inv_face_spec_rad(h, state, fArea);
  if (h->cfg.isViscous) {
    /* ref: include/spectralRadius.hpp:126-151,180-200 (Visc/FaceSpectralRadius) */
    const double rho = rho_of(h, state);
    const double gam = gamma_of(h, state);
    const double a43 = 4.0 / (3.0 * rho), gr = gam / rho;
    const double maxTerm = a43 > gr ? a43 : gr;
    const double viscTerm =
        h->cfg.nondimScaling * (mu / prandtl_of(gam) + mut / turb_prandtl(h));
    sr += fArea[3] / dist * maxTerm * viscTerm;
  }
  /* turbModel::FaceSpectralRadius; ref: include/turbulence.hpp:308-329 */
  double tsr = 0.0;
  if (h->nt > 0) {
    tsr = turb_inv_face_spec_rad(h, state, fArea, positive);
    tsr += turb_visc_spec_rad(h, state, fArea[3] / dist, mu, mut, f1);
  }
  for (int e = 0; e < neq; ++e) {
    double fc = 0.5 * fArea[3] * (fn[e] - fo[e]);
    if (e >= h->ns + 4) fc = 0.0;
    const double srd = (e < h->ns + 4 ? sr : tsr) * du[e];
    out[e] = positive ? fc + srd : fc - srd;
  }
}

static int is_physical(const orc_block *b, int i, int j, int k) {
  return i >= 0 && i < b->ni && j >= 0 && j < b->nj && k >= 0 && k < b->nk;
}
/* ref: include/boundaryConditions.hpp:287-293, src/boundaryConditions.cpp
 * GetBCSurface: is the boundary face (i,j,k) of surface type `surf` covered by
 * a connection BC? */
static int bc_is_connection(const orc_block *b, int i, int j, int k, int surf) {
  for (int s = 0; s < b->nsurf; ++s) {
    const aither_surface *sf = &b->surf[s];
    if (surface_type(sf) != surf) continue;
    int in = 0;
    if (surf <= 2) in = i == sf->imin && j >= sf->jmin && j < sf->jmax &&
                        k >= sf->kmin && k < sf->kmax;
    else if (surf <= 4) in = j == sf->jmin && i >= sf->imin && i < sf->imax &&
                             k >= sf->kmin && k < sf->kmax;
    else in = k == sf->kmin && i >= sf->imin && i < sf->imax &&
              j >= sf->jmin && j < sf->jmax;
    if (in) return sf->type == AITHER_BC_INTERBLOCK ||
                   sf->type == AITHER_BC_PERIODIC;
  }
  return 0;
}

/* ref: src/procBlock.cpp:6316-6341 (ProjC2CDist): face (ii,jj,kk) of direction d */
static double proj_c2c_dist(const orc_block *b, int d, int ii, int jj, int kk) {
  if (!b->center) return 1.0;
  const double *cu = b->center + 3 * cidx(b, ii, jj, kk);
  const double *cl = b->center + 3 * cidx(b, ii - (d == 0), jj - (d == 1), kk - (d == 2));
  const double *fa = d == 0 ? b->fAI + 4 * fidxI(b, ii, jj, kk)
                            : (d == 1 ? b->fAJ + 4 * fidxJ(b, ii, jj, kk)
                                      : b->fAK + 4 * fidxK(b, ii, jj, kk));
  return (cu[0] - cl[0]) * fa[0] + (cu[1] - cl[1]) * fa[1] + (cu[2] - cl[2]) * fa[2];
}
/* fluxJacobian.cpp OffDiagonal (:196-238): the product of one neighbour's off-diagonal block
 * with its update, by the configured method */
static void roe_flux(const orc_level *h, const double *l, const double *r,
                     const double n[3], double *flux);
static void offdiag_cell(const orc_level *h, const orc_block *b, long nb, long own,
                         const double *x, const double *fArea, int positive, double dist,
                         double *out) {
  const int neq = h->neq;
  const double *state = b->state + neq * nb, *du = x + neq * nb;
  const double mu = h->cfg.isViscous ? b->viscosity[nb] : 0.0;
  const double mut = b->eddyVisc[nb], f1 = b->f1[nb];
  if (h->cfg.invFluxJac == AITHER_JAC_APPROX_ROE) {
    /* RoeOffDiagonal; ref: src/fluxJacobian.cpp:240-296. The caller passes (.., f1, dist, ..)
     * into parameters declared (.., dist, f1, ..) (:226-228 vs :243-244): inside, `dist` holds
     * f1 and `f1` holds dist. Restated as is. */
    const double distIn = f1, f1In = dist;
    const double *diag = b->state + neq * own;
    double oldFlux[MAXEQ], newFlux[MAXEQ], su[MAXEQ];
    roe_flux(h, state, diag, fArea, oldFlux);
    update_prim_with_cons(h, state, du, su);
    if (positive) roe_flux(h, su, diag, fArea, newFlux);
    else roe_flux(h, diag, su, fArea, newFlux);
    double sr = 0.0, tsr = 0.0;
    if (h->cfg.isViscous) {
      const double rho = rho_of(h, state);
      const double gam = gamma_of(h, state);
      const double a43 = 4.0 / (3.0 * rho), gr = gam / rho;
      const double maxTerm = a43 > gr ? a43 : gr;
      const double viscTerm =
          h->cfg.nondimScaling * (mu / prandtl_of(gam) + mut / turb_prandtl(h));
      sr += fArea[3] / distIn * maxTerm * viscTerm;
      if (h->nt > 0) tsr += turb_visc_spec_rad(h, state, fArea[3] / distIn, mu, mut, f1In);
    }
    for (int e = 0; e < neq; ++e) {
      const double fc = fArea[3] * (newFlux[e] - oldFlux[e]);
      const double srd = (e < h->ns + 4 ? sr : tsr) * du[e];
      out[e] = positive ? fc + srd : fc - srd;
    }
    return;
  }
  if (h->cfg.isBlockMatrix) {
    /* RusanovBlockOffDiagonal; ref: src/fluxJacobian.cpp:164-194 */
    double J[MAXJAC], V[MAXJAC];
    rusanov_flux_jacobian(h, state, fArea, positive, J);
    if (h->cfg.isViscous) {
      approx_tsl_jacobian(h, state, mu, mut, f1, fArea, dist, positive, b->velGrad + 9 * nb, V);
      for (int q = 0; q < h->asz; ++q) J[q] = positive ? J[q] - V[q] : J[q] + V[q];
    }
    block_mult(h, J, du, out);
    return;
  }
  offdiag_scalar(h, state, du, fArea, positive, mu, mut, f1, dist, out);
}
/* ref: src/procBlock.cpp:1056-1104 (ImplicitLower) */
static void implicit_lower(const orc_level *h, const orc_block *b, int ii,
                           int jj, int kk, const double *x, double *L) {
  const int neq = h->neq;
  const long own = cidx(b, ii, jj, kk);
  double od[MAXEQ];
  for (int e = 0; e < neq; ++e) L[e] = 0.0;
  if (is_physical(b, ii - 1, jj, kk) || bc_is_connection(b, ii, jj, kk, 1)) {
    offdiag_cell(h, b, cidx(b, ii - 1, jj, kk), own, x, b->fAI + 4 * fidxI(b, ii, jj, kk), 1,
                 proj_c2c_dist(b, 0, ii, jj, kk), od);
    for (int e = 0; e < neq; ++e) L[e] += od[e];
  }
  if (is_physical(b, ii, jj - 1, kk) || bc_is_connection(b, ii, jj, kk, 3)) {
    offdiag_cell(h, b, cidx(b, ii, jj - 1, kk), own, x, b->fAJ + 4 * fidxJ(b, ii, jj, kk), 1,
                 proj_c2c_dist(b, 1, ii, jj, kk), od);
    for (int e = 0; e < neq; ++e) L[e] += od[e];
  }
  if (is_physical(b, ii, jj, kk - 1) || bc_is_connection(b, ii, jj, kk, 5)) {
    offdiag_cell(h, b, cidx(b, ii, jj, kk - 1), own, x, b->fAK + 4 * fidxK(b, ii, jj, kk), 1,
                 proj_c2c_dist(b, 2, ii, jj, kk), od);
    for (int e = 0; e < neq; ++e) L[e] += od[e];
  }
}
/* ref: src/procBlock.cpp:1106-1170 (ImplicitUpper) */
static void implicit_upper(const orc_level *h, const orc_block *b, int ii,
                           int jj, int kk, const double *x, double *U) {
  const int neq = h->neq;
  const long own = cidx(b, ii, jj, kk);
  double od[MAXEQ];
  for (int e = 0; e < neq; ++e) U[e] = 0.0;
  if (is_physical(b, ii + 1, jj, kk) || bc_is_connection(b, ii + 1, jj, kk, 2)) {
    offdiag_cell(h, b, cidx(b, ii + 1, jj, kk), own, x, b->fAI + 4 * fidxI(b, ii + 1, jj, kk),
                 0, proj_c2c_dist(b, 0, ii + 1, jj, kk), od);
    for (int e = 0; e < neq; ++e) U[e] += od[e];
  }
  if (is_physical(b, ii, jj + 1, kk) || bc_is_connection(b, ii, jj + 1, kk, 4)) {
    offdiag_cell(h, b, cidx(b, ii, jj + 1, kk), own, x, b->fAJ + 4 * fidxJ(b, ii, jj + 1, kk),
                 0, proj_c2c_dist(b, 1, ii, jj + 1, kk), od);
    for (int e = 0; e < neq; ++e) U[e] += od[e];
  }
  if (is_physical(b, ii, jj, kk + 1) || bc_is_connection(b, ii, jj, kk + 1, 6)) {
    offdiag_cell(h, b, cidx(b, ii, jj, kk + 1), own, x, b->fAK + 4 * fidxK(b, ii, jj, kk + 1),
                 0, proj_c2c_dist(b, 2, ii, jj, kk + 1), od);
    for (int e = 0; e < neq; ++e) U[e] += od[e];
  }
}

/* ref: src/linearSolver.cpp:473-507 (dplur::DPLUR) */
static void dplur_sweep(orc_level *h, orc_block *b) {
  const int neq = h->neq;
  const long np = (long)b->NI * b->NJ * b->NK * neq;
  memcpy(b->xold, b->x, sizeof(double) * np);
  for (int kk = 0; kk < b->nk; ++kk)
    for (int jj = 0; jj < b->nj; ++jj)
      for (int ii = 0; ii < b->ni; ++ii) {
        double L[MAXEQ], U[MAXEQ], rb[MAXEQ], rhs[MAXEQ];
        implicit_lower(h, b, ii, jj, kk, b->xold, L);
        implicit_upper(h, b, ii, jj, kk, b->xold, U);
        rhs_b(h, b, ii, jj, kk, rb);
        /* b + forcing + offDiagonal, offDiagonal = L - U (forcing = 0 on the finest level) */
        const double *fc = b->forcing + neq * pidx(b, ii, jj, kk);
        for (int e = 0; e < neq; ++e) rhs[e] = rb[e] + fc[e] + (L[e] - U[e]);
        diag_mult(h, b->ainv + h->asz * pidx(b, ii, jj, kk), rhs,
                  b->x + neq * cidx(b, ii, jj, kk));
      }
}

/* ref: src/linearSolver.cpp:341-383 (LUSGS_Forward) */
static void lusgs_forward(orc_level *h, orc_block *b, int sweep) {
  const int neq = h->neq;
  const long nc = (long)b->ni * b->nj * b->nk;
  for (long nn = 0; nn < nc; ++nn) {
    const int ii = b->order[3 * nn], jj = b->order[3 * nn + 1],
              kk = b->order[3 * nn + 2];
    double L[MAXEQ], U[MAXEQ], rb[MAXEQ], rhs[MAXEQ];
    implicit_lower(h, b, ii, jj, kk, b->x, L);
    if (sweep > 0 || h->cfg.matrixRequiresInit) {
      implicit_upper(h, b, ii, jj, kk, b->x, U);
      for (int e = 0; e < neq; ++e) L[e] -= U[e];
    }
    rhs_b(h, b, ii, jj, kk, rb);
    /* b = -R/theta + forcing + nm1 - mmn (forcing = 0 on the finest level; with forcing AND
     * time terms the reference adds in that order, here the time terms come first) */
    const double *fc = b->forcing + neq * pidx(b, ii, jj, kk);
    for (int e = 0; e < neq; ++e) rhs[e] = (rb[e] + fc[e]) + L[e];
    diag_mult(h, b->ainv + h->asz * pidx(b, ii, jj, kk), rhs,
              b->x + neq * cidx(b, ii, jj, kk));
  }
}
/* ref: src/linearSolver.cpp:385-428 (LUSGS_Backward) */
static void lusgs_backward(orc_level *h, orc_block *b, int sweep) {
  const int neq = h->neq;
  const long nc = (long)b->ni * b->nj * b->nk;
  for (long nn = nc - 1; nn >= 0; --nn) {
    const int ii = b->order[3 * nn], jj = b->order[3 * nn + 1],
              kk = b->order[3 * nn + 2];
    double L[MAXEQ], U[MAXEQ], rb[MAXEQ], rhs[MAXEQ], tmp[MAXEQ];
    implicit_upper(h, b, ii, jj, kk, b->x, U);
    double *xc = b->x + neq * cidx(b, ii, jj, kk);
    if (sweep > 0 || h->cfg.matrixRequiresInit) {
      implicit_lower(h, b, ii, jj, kk, b->x, L);
      rhs_b(h, b, ii, jj, kk, rb);
      const double *fc = b->forcing + neq * pidx(b, ii, jj, kk);
      for (int e = 0; e < neq; ++e) rhs[e] = ((rb[e] + fc[e]) + L[e]) - U[e];
      diag_mult(h, b->ainv + h->asz * pidx(b, ii, jj, kk), rhs, xc);
    } else {
      diag_mult(h, b->ainv + h->asz * pidx(b, ii, jj, kk), U, tmp);
      for (int e = 0; e < neq; ++e) xc[e] = xc[e] - tmp[e];
    }
  }
}

/* ref: src/linearSolver.cpp:58-109 (AXmB, Residual): f - (A x - offDiag - b) */
static void matrix_residual(orc_level *h, orc_block *b) {
  const int neq = h->neq;
  for (int kk = 0; kk < b->nk; ++kk)
    for (int jj = 0; jj < b->nj; ++jj)
      for (int ii = 0; ii < b->ni; ++ii) {
        double L[MAXEQ], U[MAXEQ], rb[MAXEQ], ax[MAXEQ];
        implicit_lower(h, b, ii, jj, kk, b->x, L);
        implicit_upper(h, b, ii, jj, kk, b->x, U);
        rhs_b(h, b, ii, jj, kk, rb);
        diag_mult(h, b->a + h->asz * pidx(b, ii, jj, kk),
                  b->x + neq * cidx(b, ii, jj, kk), ax);
        double *mr = b->mresid + neq * pidx(b, ii, jj, kk);
        for (int e = 0; e < neq; ++e)
          mr[e] = b->forcing[neq * pidx(b, ii, jj, kk) + e] - ((ax[e] - (L[e] - U[e])) - rb[e]);
      }
}

double orc_relax(orc_level *h, int sweeps) {
  /* ref: src/linearSolver.cpp:430-471 (lusgs::Relax), :509-535 (dplur::Relax),
   * src/mgSolution.cpp:198-206 (norm of the matrix residual) */
  for (int s = 0; s < sweeps; ++s) {
    swap_connections(h, 1);
    if (h->cfg.solver == AITHER_SOLVER_DPLUR) {
      for (int bb = 0; bb < h->nblk; ++bb) dplur_sweep(h, &h->blk[bb]);
    } else {
      for (int bb = 0; bb < h->nblk; ++bb) lusgs_forward(h, &h->blk[bb], s);
      swap_connections(h, 1);
      for (int bb = 0; bb < h->nblk; ++bb) lusgs_backward(h, &h->blk[bb], s);
    }
  }
  swap_connections(h, 1);
  double l2 = 0.0;
  long total = 0;
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    matrix_residual(h, b);
    /* the reference sums the squares over the ghost-padded array in storage
     * order (ghost entries are zero) */
    double sum = 0.0;
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii)
          for (int e = 0; e < h->neq; ++e) {
            const double v = b->mresid[h->neq * pidx(b, ii, jj, kk) + e];
            sum += v * v;
          }
    l2 += sum;
    total += (long)b->NI * b->NJ * b->NK * h->neq;
  }
  return l2 / (double)total;
}

void orc_update_blocks(orc_level *h, int mm, double *residL2,
                       aither_linf *linf) {
  /* ref: src/procBlock.cpp:826-871, :902-915 */
  const int neq = h->neq;
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii) {
          double *s = b->state + neq * cidx(b, ii, jj, kk);
          double ns_[MAXEQ];
          update_prim_with_cons(h, s, b->x + neq * cidx(b, ii, jj, kk), ns_);
          memcpy(s, ns_, sizeof(double) * neq);
          const double *r = b->residual + neq * pidx(b, ii, jj, kk);
          for (int e = 0; e < neq; ++e) residL2[e] += r[e] * r[e];
          for (int e = 0; e < neq; ++e) {
            if (r[e] > linf->linf) {
              linf->linf = r[e];
              linf->block = b->parentBlock;
              linf->i = ii;
              linf->j = jj;
              linf->k = kk;
              linf->eqn = e + 1;
            }
          }
        }
    /* ref: src/gridLevel.cpp:427-430: U^(n-1) <- U^n at the end of the nonlinear iterations */
    const int nl = h->cfg.nonlinearIterations > 0 ? h->cfg.nonlinearIterations : 1;
    if (h->cfg.isMultilevelTime && mm == nl - 1)
      memcpy(b->consNm1, b->consN,
             sizeof(double) * (long)b->ni * b->nj * b->nk * h->neq);
  }
}

void orc_reset_diagonal(orc_level *h) {
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    memset(b->a, 0, sizeof(double) * (long)b->ni * b->nj * b->nk * h->asz);
  }
}

void orc_store_old_solution(orc_level *h, int iter) {
  /* ref: src/mgSolution.cpp:103-114, src/procBlock.cpp:1037-1053 */
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii)
          prim_to_cons(h, b->state + h->neq * cidx(b, ii, jj, kk),
                       b->consN + h->neq * pidx(b, ii, jj, kk));
    if (h->cfg.isMultilevelTime && iter == 0)
      memcpy(b->consNm1, b->consN,
             sizeof(double) * (long)b->ni * b->nj * b->nk * h->neq);
  }
}

double orc_iterate(orc_level *h, double cfl, int mm, double *residL2,
                   aither_linf *linf) {
  /* ref: src/mgSolution.cpp:246-269, :209-244 */
  orc_get_boundary_conditions(h);
  orc_calc_residual(h);
  orc_calc_time_step(h, cfl);
  orc_invert_diagonal(h);
  orc_initialize_matrix_update(h);
  const double mr = orc_relax(h, h->cfg.matrixSweeps);
  orc_update_blocks(h, mm, residL2, linf);
  orc_reset_diagonal(h);
  return mr;
}

/* ------------------------------------------------------------------------ */
static int cmp_plane(const void *a, const void *b) {
  const int *x = (const int *)a, *y = (const int *)b;
  const int sx = x[0] + x[1] + x[2], sy = y[0] + y[1] + y[2];
  if (sx != sy) return sx - sy;
  /* within a hyperplane the order is immaterial (no intra-plane coupling) */
  if (x[2] != y[2]) return x[2] - y[2];
  if (x[1] != y[1]) return x[1] - y[1];
  return x[0] - y[0];
}

/* gridLevel::AuxillaryAndWidths -> UpdateAuxillaryVariables(phys, false): temperature
 * and viscosity of the physical cells before the first iteration (the viscous-wall
 * omega BC reads the stored viscosity); ref: src/gridLevel.cpp:433-438 */
static void init_aux(orc_level *h, orc_block *b) {
  for (int kk = 0; kk < b->nk; ++kk)
    for (int jj = 0; jj < b->nj; ++jj)
      for (int ii = 0; ii < b->ni; ++ii) {
        const long c = cidx(b, ii, jj, kk);
        b->temperature[c] = temperature_of(h, b->state + h->neq * c);
        if (h->cfg.isViscous)
          b->viscosity[c] = viscosity_of(h, b->temperature[c], b->state + h->neq * c);
      }
}

orc_level *orc_create(const aither_cfg *cfg, int nBlocks,
                      const aither_block_desc *blocks, int nConnections,
                      const aither_conn *conns) {
  orc_level *h = (orc_level *)calloc(1, sizeof(orc_level));
  h->cfg = *cfg;
  h->ns = cfg->numSpecies;
  h->nt = cfg->numTurb;
  h->neq = h->ns + 4 + h->nt;
  h->asz = cfg->isBlockMatrix ? (h->ns + 4) * (h->ns + 4) + h->nt * h->nt
                              : 1 + (h->nt > 0 ? 1 : 0); /* scalar: {flow, turb} */
  h->nblk = nBlocks;
  h->blk = (orc_block *)calloc(nBlocks, sizeof(orc_block));
  h->nconn = nConnections;
  h->conn = (aither_conn *)calloc(nConnections > 0 ? nConnections : 1,
                                  sizeof(aither_conn));
  if (nConnections > 0)
    memcpy(h->conn, conns, sizeof(aither_conn) * nConnections);
  for (int bb = 0; bb < nBlocks; ++bb) {
    orc_block *b = &h->blk[bb];
    const aither_block_desc *d = &blocks[bb];
    b->ni = d->ni;
    b->nj = d->nj;
    b->nk = d->nk;
    b->g = cfg->numGhosts;
    b->NI = b->ni + 2 * b->g;
    b->NJ = b->nj + 2 * b->g;
    b->NK = b->nk + 2 * b->g;
    b->parentBlock = d->parentBlock;
    b->nsurf = d->numSurfaces;
    b->surf = (aither_surface *)malloc(sizeof(aither_surface) * b->nsurf);
    memcpy(b->surf, d->surfaces, sizeof(aither_surface) * b->nsurf);
    const long np = (long)b->NI * b->NJ * b->NK;
    const long nc = (long)b->ni * b->nj * b->nk;
    b->state = (double *)malloc(sizeof(double) * np * h->neq);
    memcpy(b->state, d->state, sizeof(double) * np * h->neq);
    b->residual = (double *)calloc(nc * h->neq, sizeof(double));
    b->specRad = (double *)calloc(nc * 2, sizeof(double));
    b->dt = (double *)calloc(nc, sizeof(double));
    b->a = (double *)calloc(nc * h->asz, sizeof(double));
    b->ainv = (double *)calloc(nc * h->asz, sizeof(double));
    b->x = (double *)calloc(np * h->neq, sizeof(double));
    b->xold = (double *)calloc(np * h->neq, sizeof(double));
    b->consN = (double *)calloc(nc * h->neq, sizeof(double));
    b->consNm1 = (double *)calloc(nc * h->neq, sizeof(double));
    b->mresid = (double *)calloc(nc * h->neq, sizeof(double));
    b->temperature = (double *)calloc(np, sizeof(double));
    b->viscosity = (double *)calloc(np, sizeof(double));
    b->eddyVisc = (double *)calloc(np, sizeof(double));
    b->f1 = (double *)calloc(np, sizeof(double));
    b->f2 = (double *)calloc(np, sizeof(double));
    b->velGrad = (double *)calloc(np * 9, sizeof(double));
    b->tkeGrad = (double *)calloc(nc * 3, sizeof(double));
    b->omegaGrad = (double *)calloc(nc * 3, sizeof(double));
    b->pressGrad = (double *)calloc(nc * 3, sizeof(double));
    b->forcing = (double *)calloc(nc * h->neq, sizeof(double));
    b->savedX = (double *)calloc(np * h->neq, sizeof(double));
    b->vol = d->vol;
    b->fAI = d->fAreaI;
    b->fAJ = d->fAreaJ;
    b->fAK = d->fAreaK;
    b->center = d->center;
    b->cwI = d->cellWidthI;
    b->cwJ = d->cellWidthJ;
    b->cwK = d->cellWidthK;
    b->wallDist = d->wallDist;
    /* ref: src/utility.cpp:377-398 (HyperplaneReorder) */
    b->order = (int *)malloc(sizeof(int) * 3 * nc);
    b->wall = (orc_wall_vars **)calloc(b->nsurf > 0 ? b->nsurf : 1, sizeof(orc_wall_vars *));
    for (int s = 0; s < b->nsurf; ++s) {
      const aither_surface *sf = &b->surf[s];
      if (sf->type != AITHER_BC_VISCOUS_WALL) continue;
      long n = 1;
      const int ext[3] = {sf->imax - sf->imin, sf->jmax - sf->jmin, sf->kmax - sf->kmin};
      for (int q = 0; q < 3; ++q) n *= ext[q] > 0 ? ext[q] : 1;
      b->wall[s] = (orc_wall_vars *)calloc(n, sizeof(orc_wall_vars));
    }
    long q = 0;
    for (int kk = 0; kk < b->nk; ++kk)
      for (int jj = 0; jj < b->nj; ++jj)
        for (int ii = 0; ii < b->ni; ++ii) {
          b->order[3 * q] = ii;
          b->order[3 * q + 1] = jj;
          b->order[3 * q + 2] = kk;
          ++q;
        }
    qsort(b->order, nc, 3 * sizeof(int), cmp_plane);
    init_aux(h, b);
  }
  return h;
}

void orc_destroy(orc_level *h) {
  if (!h) return;
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    free(b->surf); free(b->state); free(b->residual); free(b->specRad);
    free(b->dt); free(b->a); free(b->ainv); free(b->x); free(b->xold);
    free(b->consN); free(b->consNm1); free(b->mresid); free(b->temperature);
    free(b->viscosity); free(b->order);
    free(b->eddyVisc); free(b->f1); free(b->f2); free(b->velGrad);
    free(b->tkeGrad); free(b->omegaGrad); free(b->pressGrad);
    free(b->forcing); free(b->savedX); free(b->toCoarse); free(b->volFac); free(b->prolong);
    if (b->wall) {
      for (int s = 0; s < b->nsurf; ++s) free(b->wall[s]);
      free(b->wall);
    }
  }
  free(h->blk);
  free(h->conn);
  free(h);
}

long long orc_field_size(orc_level *h, int blk, int field) {
  const orc_block *b = &h->blk[blk];
  const long long np = (long long)b->NI * b->NJ * b->NK;
  const long long nc = (long long)b->ni * b->nj * b->nk;
  switch (field) {
    case AITHER_FIELD_STATE: return np * h->neq;
    case AITHER_FIELD_RESIDUAL: return nc * h->neq;
    case AITHER_FIELD_SPEC_RADIUS: return nc * 2;
    case AITHER_FIELD_DT: return nc;
    case AITHER_FIELD_DIAG: return nc * h->asz;
    case AITHER_FIELD_DIAG_INV: return nc * h->asz;
    case AITHER_FIELD_UPDATE: return np * h->neq;
    case AITHER_FIELD_CONS_N: return nc * h->neq;
    case AITHER_FIELD_MATRIX_RESID: return nc * h->neq;
    case AITHER_FIELD_TEMPERATURE: return np;
    case AITHER_FIELD_VISCOSITY: return np;
    case AITHER_FIELD_CONS_NM1: return nc * h->neq;
    case AITHER_FIELD_EDDY_VISCOSITY: return np;
    case AITHER_FIELD_F1: return np;
    case AITHER_FIELD_F2: return np;
    case AITHER_FIELD_VELOCITY_GRAD: return np * 9;
    case AITHER_FIELD_TKE_GRAD: return nc * 3;
    case AITHER_FIELD_OMEGA_GRAD: return nc * 3;
    case AITHER_FIELD_PRESSURE_GRAD: return nc * 3;
  }
  return 0;
}

/* wallVars of one viscous-wall surface, 12 doubles per face in the order the reference's wall
 * function file uses (y+, shear stress x / y / z, heat flux, wall temperature, wall eddy
 * viscosity, wall viscosity, wall density, friction velocity, k, omega; include/wallData.hpp:40-57),
 * faces i fastest, then j, then k over the surface's cell range. Returns the number of faces, 0 for
 * a surface that keeps none. */
long long orc_get_wall_data(orc_level *h, int blk, int surface, double *dst) {
  const orc_block *b = &h->blk[blk];
  if (surface < 0 || surface >= b->nsurf || !b->wall || !b->wall[surface]) return 0;
  const aither_surface *sf = &b->surf[surface];
  const int ext[3] = {sf->imax - sf->imin, sf->jmax - sf->jmin, sf->kmax - sf->kmin};
  long n = 1;
  for (int q = 0; q < 3; ++q) n *= ext[q] > 0 ? ext[q] : 1;
  for (long f = 0; f < n && dst; ++f) {
    const orc_wall_vars *w = &b->wall[surface][f];
    double *o = dst + 12 * f;
    o[0] = w->yplus;
    o[1] = w->shearStress[0];
    o[2] = w->shearStress[1];
    o[3] = w->shearStress[2];
    o[4] = w->heatFlux;
    o[5] = w->temperature;
    o[6] = w->turbEddyVisc;
    o[7] = w->viscosity;
    o[8] = w->density;
    o[9] = w->frictionVelocity;
    o[10] = w->tke;
    o[11] = w->sdr;
  }
  return n;
}

void orc_get_field(orc_level *h, int blk, int field, double *dst) {
  const orc_block *b = &h->blk[blk];
  const double *src = NULL;
  switch (field) {
    case AITHER_FIELD_STATE: src = b->state; break;
    case AITHER_FIELD_RESIDUAL: src = b->residual; break;
    case AITHER_FIELD_SPEC_RADIUS: src = b->specRad; break;
    case AITHER_FIELD_DT: src = b->dt; break;
    case AITHER_FIELD_DIAG: src = b->a; break;
    case AITHER_FIELD_DIAG_INV: src = b->ainv; break;
    case AITHER_FIELD_UPDATE: src = b->x; break;
    case AITHER_FIELD_CONS_N: src = b->consN; break;
    case AITHER_FIELD_MATRIX_RESID: src = b->mresid; break;
    case AITHER_FIELD_TEMPERATURE: src = b->temperature; break;
    case AITHER_FIELD_VISCOSITY: src = b->viscosity; break;
    case AITHER_FIELD_CONS_NM1: src = b->consNm1; break;
    case AITHER_FIELD_EDDY_VISCOSITY: src = b->eddyVisc; break;
    case AITHER_FIELD_F1: src = b->f1; break;
    case AITHER_FIELD_F2: src = b->f2; break;
    case AITHER_FIELD_VELOCITY_GRAD: src = b->velGrad; break;
    case AITHER_FIELD_TKE_GRAD: src = b->tkeGrad; break;
    case AITHER_FIELD_OMEGA_GRAD: src = b->omegaGrad; break;
    case AITHER_FIELD_PRESSURE_GRAD: src = b->pressGrad; break;
  }
  if (src) memcpy(dst, src, sizeof(double) * orc_field_size(h, blk, field));
}

void orc_set_state(orc_level *h, int blk, const double *stateAoS) {
  memcpy(h->blk[blk].state, stateAoS,
         sizeof(double) * orc_field_size(h, blk, AITHER_FIELD_STATE));
  init_aux(h, &h->blk[blk]);
}

/* ------------------------------------------------------------------------ */
/* point-function exports                                                     */
static void level_from_cfg(orc_level *h, const aither_cfg *cfg) {
  memset(h, 0, sizeof(*h));
  h->cfg = *cfg;
  h->ns = cfg->numSpecies;
  h->nt = cfg->numTurb;
  h->neq = h->ns + 4 + h->nt;
  h->asz = cfg->isBlockMatrix ? (h->ns + 4) * (h->ns + 4) + h->nt * h->nt
                              : 1 + (h->nt > 0 ? 1 : 0);
}
void orc_inviscid_flux(const aither_cfg *cfg, const double *left,
                       const double *right, const double nrm[3], double *flux) {
  orc_level h;
  level_from_cfg(&h, cfg);
  inviscid_flux(&h, left, right, nrm, flux);
}
void orc_ghost_state(const aither_cfg *cfg, const double *interior, int bcType,
                     const double areaUnit[3], int surfType, int tag, int layer,
                     double *ghost) {
  orc_level h;
  level_from_cfg(&h, cfg);
  ghost_state(&h, interior, bcType, areaUnit, surfType, tag, layer, 0.0, 0.0, ghost, NULL);
}
void orc_offdiag_scalar(const aither_cfg *cfg, const double *stateNb,
                        const double *duNb, const double fArea[4], int positive,
                        double *out) {
  orc_level h;
  level_from_cfg(&h, cfg);
  offdiag_scalar(&h, stateNb, duNb, fArea, positive, 0.0, 0.0, 0.0, 1.0, out);
}

/* turbulence point functions (tests/test_physics_host.py) */
void orc_eddy_visc(const aither_cfg *cfg, const double *state, const double vg[9],
                   const double kg[3], const double wg[3], double mu, double wallDist,
                   double out[3]) {
  orc_level h;
  level_from_cfg(&h, cfg);
  eddy_visc_and_blending(&h, state, vg, kg, wg, mu, wallDist, &out[0], &out[1], &out[2]);
}
void orc_turb_source(const aither_cfg *cfg, const double *state, const double vg[9],
                     const double kg[3], const double wg[3], double mut, double f1,
                     double src[2]) {
  orc_level h;
  double beta;
  level_from_cfg(&h, cfg);
  calc_turb_src(&h, state, vg, kg, wg, mut, f1, src, &beta);
}
void orc_offdiag_scalar_visc(const aither_cfg *cfg, const double *stateNb, const double *duNb,
                             const double fArea[4], int positive, double mu, double mut,
                             double f1, double dist, double *out) {
  orc_level h;
  level_from_cfg(&h, cfg);
  offdiag_scalar(&h, stateNb, duNb, fArea, positive, mu, mut, f1, dist, out);
}
void orc_ghost_state_visc(const aither_cfg *cfg, const double *interior, int bcType,
                          const double areaUnit[3], int surfType, int tag, int layer,
                          double wallDist, double nuW, double *ghost) {
  orc_level h;
  level_from_cfg(&h, cfg);
  ghost_state(&h, interior, bcType, areaUnit, surfType, tag, layer, wallDist, nuW, ghost, NULL);
}

/* transport of a mixture state (tests/test_physics_host.py): {viscosity, effective conductivity} */
/* ------------------------------------------------------------------------ */
/* multigrid transfer operators between two levels (each an orc_level of its own;
 * ref: src/gridLevel.cpp:538-611, include/procBlock.hpp:637-690,
 * include/gridLevel.hpp:159-214, include/utility.hpp:186-372). The cycle itself
 * (mgSolution::CycleAtLevel, src/mgSolution.cpp:160-207) is composed from these and
 * the per-level phases by the caller (tests/oracle.py).                      */
void orc_set_transfer(orc_level *h, int blk, const int *toCoarse, const double *volFac,
                      const double *prolong) {
  orc_block *b = &h->blk[blk];
  const long nc = (long)b->ni * b->nj * b->nk;
  free(b->toCoarse); free(b->volFac); free(b->prolong);
  b->toCoarse = (int *)malloc(sizeof(int) * 3 * nc);
  b->volFac = (double *)malloc(sizeof(double) * nc);
  b->prolong = (double *)malloc(sizeof(double) * 7 * nc);
  memcpy(b->toCoarse, toCoarse, sizeof(int) * 3 * nc);
  memcpy(b->volFac, volFac, sizeof(double) * nc);
  memcpy(b->prolong, prolong, sizeof(double) * 7 * nc);
}

/* gridLevel::Restriction; `cfl` = inp.CFL() of the current iteration */
void orc_mg_restrict(orc_level *fine, orc_level *coarse, int mm, double cfl) {
  const int neq = fine->neq;
  for (int bb = 0; bb < fine->nblk; ++bb) {
    const orc_block *f = &fine->blk[bb];
    orc_block *c = &coarse->blk[bb];
    /* volume-weighted state (BlockRestriction zeroes the coarse array, ghosts included) */
    memset(c->state, 0, sizeof(double) * (long)c->NI * c->NJ * c->NK * neq);
    for (int kk = 0; kk < f->nk; ++kk)
      for (int jj = 0; jj < f->nj; ++jj)
        for (int ii = 0; ii < f->ni; ++ii) {
          const long pf = pidx(f, ii, jj, kk);
          const int *ci = f->toCoarse + 3 * pf;
          double *cs = c->state + neq * cidx(c, ci[0], ci[1], ci[2]);
          const double *fs = f->state + neq * cidx(f, ii, jj, kk);
          for (int e = 0; e < neq; ++e) cs[e] = cs[e] + f->volFac[pf] * fs[e];
        }
    if (mm == 0) orc_store_old_solution(coarse, 1);
  }
  orc_get_boundary_conditions(coarse);
  orc_calc_residual(coarse);
  orc_calc_time_step(coarse, cfl);
  orc_invert_diagonal(coarse);
  /* linearSolver::Restriction: volume-weighted update, then its ghost swap */
  for (int bb = 0; bb < fine->nblk; ++bb) {
    const orc_block *f = &fine->blk[bb];
    orc_block *c = &coarse->blk[bb];
    memset(c->x, 0, sizeof(double) * (long)c->NI * c->NJ * c->NK * neq);
    for (int kk = 0; kk < f->nk; ++kk)
      for (int jj = 0; jj < f->nj; ++jj)
        for (int ii = 0; ii < f->ni; ++ii) {
          const long pf = pidx(f, ii, jj, kk);
          const int *ci = f->toCoarse + 3 * pf;
          double *cx = c->x + neq * cidx(c, ci[0], ci[1], ci[2]);
          const double *fx = f->x + neq * cidx(f, ii, jj, kk);
          for (int e = 0; e < neq; ++e) cx[e] = cx[e] + f->volFac[pf] * fx[e];
        }
  }
  swap_connections(coarse, 1);
  /* forcing = (A x - b) of the coarse level with the restricted update and state + the fine
   * level's matrix residual summed over the fine cells of every coarse cell */
  for (int bb = 0; bb < fine->nblk; ++bb) {
    const orc_block *f = &fine->blk[bb];
    orc_block *c = &coarse->blk[bb];
    const long ncc = (long)c->ni * c->nj * c->nk;
    memset(c->forcing, 0, sizeof(double) * ncc * neq);
    for (int kk = 0; kk < f->nk; ++kk)
      for (int jj = 0; jj < f->nj; ++jj)
        for (int ii = 0; ii < f->ni; ++ii) {
          const long pf = pidx(f, ii, jj, kk);
          const int *ci = f->toCoarse + 3 * pf;
          double *cf = c->forcing + neq * pidx(c, ci[0], ci[1], ci[2]);
          for (int e = 0; e < neq; ++e) cf[e] = cf[e] + f->mresid[neq * pf + e];
        }
    for (int kk = 0; kk < c->nk; ++kk)
      for (int jj = 0; jj < c->nj; ++jj)
        for (int ii = 0; ii < c->ni; ++ii) {
          double L[MAXEQ], U[MAXEQ], rb[MAXEQ], ax[MAXEQ];
          implicit_lower(coarse, c, ii, jj, kk, c->x, L);
          implicit_upper(coarse, c, ii, jj, kk, c->x, U);
          rhs_b(coarse, c, ii, jj, kk, rb);
          diag_mult(coarse, c->a + coarse->asz * pidx(c, ii, jj, kk),
                    c->x + neq * cidx(c, ii, jj, kk), ax);
          double *cf = c->forcing + neq * pidx(c, ii, jj, kk);
          for (int e = 0; e < neq; ++e) cf[e] = ((ax[e] - (L[e] - U[e])) - rb[e]) + cf[e];
        }
  }
}

void orc_mg_save_update(orc_level *h) {
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    memcpy(b->savedX, b->x, sizeof(double) * (long)b->NI * b->NJ * b->NK * h->neq);
  }
}
void orc_mg_subtract_saved(orc_level *h) { /* linearSolver::SubtractFromUpdate */
  for (int bb = 0; bb < h->nblk; ++bb) {
    orc_block *b = &h->blk[bb];
    const long n = (long)b->NI * b->NJ * b->NK * h->neq;
    for (long q = 0; q < n; ++q) b->x[q] -= b->savedX[q];
  }
}

/* gridLevel::Prolongation: coarse update -> node values of the coarse block
 * (ConvertCellToNode(coarse, ignoreEdge = true, ignoreGhosts = true)) -> trilinear
 * interpolation at every fine cell centre -> added to the fine update */
void orc_mg_prolong(orc_level *coarse, orc_level *fine) {
  const int neq = fine->neq;
  for (int bb = 0; bb < fine->nblk; ++bb) {
    const orc_block *c = &coarse->blk[bb];
    orc_block *f = &fine->blk[bb];
    const int n0 = c->ni + 1, n1 = c->nj + 1, n2 = c->nk + 1;
    double *node = (double *)calloc((size_t)n0 * n1 * n2 * neq, sizeof(double));
#define ND(i, j, k) (node + neq * ((long)(i) + (long)n0 * ((j) + (long)n1 * (k))))
    const int off[8][3] = {{0, 0, 0}, {0, 1, 0}, {0, 1, 1}, {0, 0, 1},
                           {1, 0, 0}, {1, 1, 0}, {1, 1, 1}, {1, 0, 1}};
    for (int kk = 0; kk < c->nk; ++kk)
      for (int jj = 0; jj < c->nj; ++jj)
        for (int ii = 0; ii < c->ni; ++ii) {
          const double *cx = c->x + neq * cidx(c, ii, jj, kk);
          for (int q = 0; q < 8; ++q) {
            double *nd = ND(ii + off[q][0], jj + off[q][1], kk + off[q][2]);
            for (int e = 0; e < neq; ++e) nd[e] += cx[e];
          }
        }
    for (int kk = 0; kk < n2; ++kk)
      for (int jj = 0; jj < n1; ++jj)
        for (int ii = 0; ii < n0; ++ii) {
          const int ei = ii == 0 || ii == n0 - 1, ej = jj == 0 || jj == n1 - 1,
                    ek = kk == 0 || kk == n2 - 1;
          /* AtInteriorCorner: 1; AtInteriorEdge: 1/2; else 1/8 (no ghost layers) */
          const double fac = (ei && ej && ek) ? 1.0 : ((ei + ej + ek == 2) ? 1.0 / 2.0 : 1.0 / 8.0);
          double *nd = ND(ii, jj, kk);
          for (int e = 0; e < neq; ++e) nd[e] *= fac;
        }
    for (int kk = 0; kk < f->nk; ++kk)
      for (int jj = 0; jj < f->nj; ++jj)
        for (int ii = 0; ii < f->ni; ++ii) {
          const long pf = pidx(f, ii, jj, kk);
          const int *ci = f->toCoarse + 3 * pf;
          const double *co = f->prolong + 7 * pf;
          const double *d0 = ND(ci[0], ci[1], ci[2]), *d1 = ND(ci[0] + 1, ci[1], ci[2]),
                       *d2 = ND(ci[0], ci[1] + 1, ci[2]), *d3 = ND(ci[0] + 1, ci[1] + 1, ci[2]),
                       *d4 = ND(ci[0], ci[1], ci[2] + 1), *d5 = ND(ci[0] + 1, ci[1], ci[2] + 1),
                       *d6 = ND(ci[0], ci[1] + 1, ci[2] + 1),
                       *d7 = ND(ci[0] + 1, ci[1] + 1, ci[2] + 1);
          double *fx = f->x + neq * cidx(f, ii, jj, kk);
          for (int e = 0; e < neq; ++e) {
            /* TrilinearInterp: LinearInterp(a, b, c) = (1 - c) a + c b */
            const double d04 = (1.0 - co[0]) * d0[e] + co[0] * d4[e];
            const double d15 = (1.0 - co[1]) * d1[e] + co[1] * d5[e];
            const double d26 = (1.0 - co[2]) * d2[e] + co[2] * d6[e];
            const double d37 = (1.0 - co[3]) * d3[e] + co[3] * d7[e];
            const double d0415 = (1.0 - co[4]) * d04 + co[4] * d15;
            const double d2637 = (1.0 - co[5]) * d26 + co[5] * d37;
            fx[e] += (1.0 - co[6]) * d0415 + co[6] * d2637;
          }
        }
#undef ND
    free(node);
  }
}

/* ghost state of a non-reflecting inlet / pressure outlet; extra = {dt, stateN[neq],
 * pressGrad[3], velGrad[9], avgMach, maxMach} (tests/test_physics_host.py) */
void orc_ghost_state_nonreflecting(const aither_cfg *cfg, const double *interior, int bcType,
                                   const double areaUnit[3], int surfType, int tag, int layer,
                                   const double *extra, double *ghost) {
  orc_level h;
  level_from_cfg(&h, cfg);
  orc_bc_extra ex;
  memset(&ex, 0, sizeof(ex));
  ex.dt = extra[0];
  for (int e = 0; e < h.neq; ++e) ex.stateN[e] = extra[1 + e];
  for (int q = 0; q < 3; ++q) ex.pressGrad[q] = extra[1 + h.neq + q];
  for (int q = 0; q < 9; ++q) ex.velGrad[q] = extra[4 + h.neq + q];
  ex.avgMach = extra[13 + h.neq];
  ex.maxMach = extra[14 + h.neq];
  g_bc_extra = &ex;
  ghost_state(&h, interior, bcType, areaUnit, surfType, tag, layer, 0.0, 0.0, ghost, NULL);
  g_bc_extra = NULL;
}

/* wallLaw::AdiabaticBCs (mode 0) / HeatFluxBCs (1) / IsothermalBCs (2) for the boundary state
 * `tag`; out = {yplus, tau x y z, heatFlux, viscosity, eddy viscosity, density, temperature,
 * tke, sdr} (tests/test_physics_host.py) */
void orc_wall_law(const aither_cfg *cfg, int mode, int tag, const double *state,
                  double wallDist, const double area[3], int isLower, double out[11]) {
  orc_level h;
  level_from_cfg(&h, cfg);
  orc_wall_vars wv;
  memset(&wv, 0, sizeof(wv));
  wall_law_eval(&h, bc_data(&h, tag), mode, state, wallDist, area, isLower, &wv);
  out[0] = wv.yplus;
  out[1] = wv.shearStress[0];
  out[2] = wv.shearStress[1];
  out[3] = wv.shearStress[2];
  out[4] = wv.heatFlux;
  out[5] = wv.viscosity;
  out[6] = wv.turbEddyVisc;
  out[7] = wv.density;
  out[8] = wv.temperature;
  out[9] = wv.tke;
  out[10] = wv.sdr;
}

void orc_mixture_transport(const aither_cfg *cfg, const double *state, double out[2]) {
  orc_level h;
  level_from_cfg(&h, cfg);
  const double t = temperature_of(&h, state);
  out[0] = viscosity_of(&h, t, state);
  out[1] = eff_conductivity(&h, t, state);
}
