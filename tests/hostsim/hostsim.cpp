// hostsim.cpp -- TEST-ONLY host build of the device point functions in
// aither_b200/csrc/physics.cuh, so they can be checked against the oracle on a machine without a
// GPU. Never part of libaither_b200.so; built on demand by tests/test_physics_host.py.
#include "../../aither_b200/csrc/physics.cuh"
#include "../../aither_b200/csrc/turbulence.cuh"
#include "../../aither_b200/csrc/walllaw.cuh"

using namespace aither;

static Gas GasFromCfg(const aither_cfg *c) {
  Gas g;
  for (int s = 0; s < AITHER_MAX_SPECIES; ++s) {
    g.R[s] = c->gasConstant[s];
    g.n[s] = c->n[s];
    g.hf[s] = c->hf[s];
  }
  GasFinalize(&g);
  return g;
}
static const aither_bc_state *Find(const aither_cfg *c, int tag) {
  for (int b = 0; b < c->numBCStates; ++b)
    if (c->bcStates[b].tag == tag) return &c->bcStates[b];
  return &c->bcStates[0];
}

extern "C" {
void hs_muscl(const double *u2, const double *u1, const double *d1, double kappa, int lim,
              double w2, double w1, double wd, double *face) {
  if (lim == 0) Muscl<5, 0>(u2, u1, d1, kappa, w2, w1, wd, face);
  else if (lim == 1) Muscl<5, 1>(u2, u1, d1, kappa, w2, w1, wd, face);
  else Muscl<5, 2>(u2, u1, d1, kappa, w2, w1, wd, face);
}
void hs_weno(const double *y, const double *w, int n, int wenoz, double *face) {
  // y: 5 x n row-major (stencil-major)
  const WenoGeom g = WenoSetup(w);
  for (int e = 0; e < n; ++e) {
    face[e] = wenoz ? Weno1<true>(g, y[e], y[n + e], y[2 * n + e], y[3 * n + e], y[4 * n + e])
                    : Weno1<false>(g, y[e], y[n + e], y[2 * n + e], y[3 * n + e], y[4 * n + e]);
  }
}
void hs_inviscid_flux(const aither_cfg *c, const double *l, const double *r, const double *n,
                      double *f) {
  const Gas g = GasFromCfg(c);
  if (c->invFlux == AITHER_FLUX_ROE) RoeFlux<1, 0>(g, l, r, n, f);
  else AusmFlux<1, 0>(g, l, r, n, f);
}
void hs_ghost_state(const aither_cfg *c, const double *interior, int bcType, const double *area,
                    int surf, int tag, int layer, double *ghost) {
  const Gas g = GasFromCfg(c);
  GhostState<1, 0>(g, interior, bcType, area, surf, *Find(c, tag), layer, ghost);
}
void hs_offdiag_scalar(const aither_cfg *c, const double *s, const double *du, const double *fa,
                       int positive, double *out) {
  const Gas g = GasFromCfg(c);
  OffDiagScalar<1, 0>(g, s, du, fa, positive != 0, out);
}
void hs_update_prim(const aither_cfg *c, const double *s, const double *du, double *out) {
  const Gas g = GasFromCfg(c);
  UpdatePrimWithCons<1, 0>(g, s, du, out);
}
}

// ---- restructured twins used by the marching kernels ----------------------------------------
extern "C" {
void hs_roe_flux_fast(const aither_cfg *c, const double *l, const double *r, const double *n,
                      double *f) {
  const Gas g = GasFromCfg(c);
  RoeFluxFast<1, 0>(g, l, r, n, f);
}
// the off-diagonal product of one neighbour, through MakeIngr + OffDiagFromIngr
void hs_offdiag_fast(const aither_cfg *c, const double *s, const double *du, const double *fa,
                     int positive, double *out) {
  const Gas g = GasFromCfg(c);
  constexpr int neq = 5;
  double ing[Ingr<1, 0>::n];
  MakeIngr<1, 0>(g, s, du, &ing[neq], &ing[neq + 1], &ing[2 * neq + 2], &ing[3 * neq + 2]);
  for (int e = 0; e < neq; ++e) {
    ing[e] = s[e];
    ing[neq + 2 + e] = du[e];
    out[e] = 0.0;
  }
  auto ld = [&](int q) { return ing[q]; };
  OffDiagFromIngr<1, 0>(ld, fa, positive != 0, out);
}
}

// ---- RANS point functions (turbulence.cuh) and the NT = 2 variants ---------------------------
static Transport TransportFromCfg(const aither_cfg *c) {
  Transport t;
  t.tRef = c->tRef; t.muRef = c->muMixRef; t.kRef = c->kMixRef;
  for (int q = 0; q < AITHER_MAX_SPECIES; ++q) {
    t.viscC1[q] = c->suthViscC1[q]; t.viscS[q] = c->suthViscS[q];
    t.condC1[q] = c->suthCondC1[q]; t.condS[q] = c->suthCondS[q];
    t.molarMass[q] = c->molarMass[q];
  }
  t.schmidt = c->schmidt;
  t.scaling = c->nondimScaling; t.turbModel = c->turbModel;
  return t;
}
extern "C" {
void hs_eddy_visc(const aither_cfg *c, const double *s, const double *vg, const double *kg,
                  const double *wg, double mu, double wallDist, double *out) {
  EddyViscAndBlending(c->turbModel, c->nondimScaling, s[0], s[5], s[6], vg, kg, wg, mu, wallDist,
                      &out[0], &out[1], &out[2]);
}
void hs_turb_source(const aither_cfg *c, const double *s, const double *vg, const double *kg,
                    const double *wg, double mut, double f1, double *src) {
  TurbSource(c->turbModel, c->nondimScaling, s[0], s[5], s[6], vg, kg, wg, mut, f1, src);
}
void hs_offdiag_scalar_rans(const aither_cfg *c, const double *s, const double *du,
                            const double *fa, int positive, double mu, double mut, double f1,
                            double dist, double *out) {
  const Gas g = GasFromCfg(c);
  const Transport tr = TransportFromCfg(c);
  const double length = fa[3] / dist;
  const double extra = length * ViscSpecFactor(tr, s[0], Gamma<1>(g, s), mu, mut);
  const double extraT =
      length * TurbViscSpecFactor(tr.turbModel, tr.scaling, s[0], s[5], s[6], mu, mut, f1);
  OffDiagScalar<1, 2>(g, s, du, fa, positive != 0, out, extra, extraT);
}
void hs_ghost_state_rans(const aither_cfg *c, const double *interior, int bcType,
                         const double *area, int surf, int tag, int layer, double *ghost) {
  const Gas g = GasFromCfg(c);
  const Transport tr = TransportFromCfg(c);
  GhostState<1, 2>(g, interior, bcType, area, surf, *Find(c, tag), layer, ghost, &tr);
}
void hs_ghost_state_nonreflecting(const aither_cfg *c, const double *interior, int bcType,
                                  const double *area, int surf, int tag, int layer,
                                  const double *extra, double *ghost) {
  const Gas g = GasFromCfg(c);
  BcExtra ex;
  ex.dt = extra[0];
  for (int e = 0; e < 5; ++e) ex.stateN[e] = extra[1 + e];
  for (int q = 0; q < 3; ++q) ex.pressGrad[q] = extra[6 + q];
  for (int q = 0; q < 9; ++q) ex.velGrad[q] = extra[9 + q];
  ex.avgMach = extra[18];
  ex.maxMach = extra[19];
  GhostState<1, 0>(g, interior, bcType, area, surf, *Find(c, tag), layer, ghost, nullptr, &ex);
}
void hs_wall_law(const aither_cfg *c, int mode, int tag, const double *state, double wallDist,
                 const double *area, int isLower, double *out) {
  const Gas g = GasFromCfg(c);
  const Transport tr = TransportFromCfg(c);
  WallVars wv;
  WallLawEval<1, 2>(g, tr, *Find(c, tag), mode, state, wallDist, area, isLower != 0, wv);
  out[0] = wv.yplus;
  out[1] = wv.tau[0];
  out[2] = wv.tau[1];
  out[3] = wv.tau[2];
  out[4] = wv.heatFlux;
  out[5] = wv.mu;
  out[6] = wv.mut;
  out[7] = wv.rho;
  out[8] = wv.t;
  out[9] = wv.tke;
  out[10] = wv.sdr;
}
void hs_inviscid_flux_rans(const aither_cfg *c, const double *l, const double *r, const double *n,
                           int fast, double *f) {
  const Gas g = GasFromCfg(c);
  if (c->invFlux == AITHER_FLUX_ROE) {
    if (fast) RoeFluxFast<1, 2>(g, l, r, n, f);
    else RoeFlux<1, 2>(g, l, r, n, f);
  } else {
    if (fast) InviscidFluxFast<1, 2, AITHER_FLUX_AUSM>(g, l, r, n, f);
    else AusmFlux<1, 2>(g, l, r, n, f);
  }
}
}

// ---- three-species variants --------------------------------------------------------------------
extern "C" {
void hs_mixture_transport3(const aither_cfg *c, const double *s, double *out) {
  const Gas g = GasFromCfg(c);
  const Transport tr = TransportFromCfg(c);
  const double t = Temperature<3>(g, s);
  out[0] = MixtureViscosity<3>(tr, t, s);
  out[1] = MixtureEffConductivity<3>(tr, t, s);
}
void hs_inviscid_flux3(const aither_cfg *c, const double *l, const double *r, const double *n,
                       int fast, double *f) {
  const Gas g = GasFromCfg(c);
  if (c->invFlux == AITHER_FLUX_ROE) {
    if (fast) RoeFluxFast<3, 0>(g, l, r, n, f);
    else RoeFlux<3, 0>(g, l, r, n, f);
  } else {
    if (fast) InviscidFluxFast<3, 0, AITHER_FLUX_AUSM>(g, l, r, n, f);
    else AusmFlux<3, 0>(g, l, r, n, f);
  }
}
void hs_update_prim3(const aither_cfg *c, const double *s, const double *du, double *out) {
  const Gas g = GasFromCfg(c);
  UpdatePrimWithCons<3, 0>(g, s, du, out);
}
void hs_offdiag_scalar3(const aither_cfg *c, const double *s, const double *du, const double *fa,
                        int positive, double *out) {
  const Gas g = GasFromCfg(c);
  OffDiagScalar<3, 0>(g, s, du, fa, positive != 0, out);
}
}
