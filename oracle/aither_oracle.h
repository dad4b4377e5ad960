/* aither_oracle.h -- plain-C CPU restatement of the reference hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (aither_b200/) never links, imports or executes it.
 *
 * It restates, loop for loop and in the reference's own accumulation order
 * (SURVEY.md appendix C), the per-iteration path of mnucci32/aither v0.10.0:
 * mgSolution::Iterate (src/mgSolution.cpp:246-269). Every function cites the
 * reference file:line it follows. Parity pinning: tests/test_oracle_*.py check
 * it against arrays dumped from the unmodified reference (oracle/_ref, built
 * by oracle/Makefile) and against the committed fixtures in tests/golden/,
 * which in turn reproduce the reference's regression goldens
 * (testCases/regressionTests.py:241-242, :333-334).
 *
 * Data layouts are the reference's (array-of-structs, i fastest, ghost padded)
 * and the configuration / block / connection PODs are the ones of the C ABI
 * (include/aither_gpu.h), so the same inputs feed the oracle and the GPU path.
 */
#ifndef AITHER_ORACLE_H
#define AITHER_ORACLE_H
#include "../include/aither_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_level orc_level;

orc_level *orc_create(const aither_cfg *cfg, int nBlocks,
                      const aither_block_desc *blocks, int nConnections,
                      const aither_conn *conns);
void orc_destroy(orc_level *h);

void orc_store_old_solution(orc_level *h, int iter);
void orc_get_boundary_conditions(orc_level *h);
void orc_calc_residual(orc_level *h);
void orc_calc_time_step(orc_level *h, double cfl);
void orc_invert_diagonal(orc_level *h);
void orc_initialize_matrix_update(orc_level *h);
double orc_relax(orc_level *h, int sweeps);
void orc_update_blocks(orc_level *h, int mm, double *residL2, aither_linf *linf);
void orc_reset_diagonal(orc_level *h);
double orc_iterate(orc_level *h, double cfl, int mm, double *residL2,
                   aither_linf *linf);

long long orc_field_size(orc_level *h, int blk, int field);
void orc_get_field(orc_level *h, int blk, int field, double *dst);
/* wallVars of a viscous-wall surface (12 doubles per face, the reference's wallData order);
 * returns the number of faces (dst may be NULL to ask for it) */
long long orc_get_wall_data(orc_level *h, int blk, int surface, double *dst);
void orc_set_state(orc_level *h, int blk, const double *stateAoS);

/* point functions, exported for unit tests against the device functions */
void orc_muscl(const double *uw2, const double *uw1, const double *dw1, int n,
               double kappa, int limiter, double w2, double w1, double wd,
               double *face);
void orc_weno(const double *u[5], const double w[5], int n, int isWenoZ,
              double *face);
void orc_inviscid_flux(const aither_cfg *cfg, const double *left,
                       const double *right, const double nrm[3], double *flux);
void orc_ghost_state(const aither_cfg *cfg, const double *interior, int bcType,
                     const double areaUnit[3], int surfType, int tag, int layer,
                     double *ghost);
void orc_offdiag_scalar(const aither_cfg *cfg, const double *stateNb,
                        const double *duNb, const double fArea[4], int positive,
                        double *out);

/* turbulence / viscous variants */
void orc_eddy_visc(const aither_cfg *cfg, const double *state, const double vg[9],
                   const double kg[3], const double wg[3], double mu, double wallDist,
                   double out[3] /* mut, f1, f2 */);
void orc_turb_source(const aither_cfg *cfg, const double *state, const double vg[9],
                     const double kg[3], const double wg[3], double mut, double f1,
                     double src[2]);
void orc_offdiag_scalar_visc(const aither_cfg *cfg, const double *stateNb, const double *duNb,
                             const double fArea[4], int positive, double mu, double mut,
                             double f1, double dist, double *out);
void orc_ghost_state_visc(const aither_cfg *cfg, const double *interior, int bcType,
                          const double areaUnit[3], int surfType, int tag, int layer,
                          double wallDist, double nuW, double *ghost);

/* multigrid transfer operators between two levels (gridLevel::Restriction / Prolongation,
 * linearSolver::SubtractFromUpdate); the cycle is composed by the caller (tests/oracle.py) */
void orc_set_transfer(orc_level *h, int blk, const int *toCoarse, const double *volFac,
                      const double *prolong);
void orc_mg_restrict(orc_level *fine, orc_level *coarse, int mm, double cfl);
void orc_mg_save_update(orc_level *h);
void orc_mg_subtract_saved(orc_level *h);
void orc_mg_prolong(orc_level *coarse, orc_level *fine);

void orc_ghost_state_nonreflecting(const aither_cfg *cfg, const double *interior, int bcType,
                                   const double areaUnit[3], int surfType, int tag, int layer,
                                   const double *extra, double *ghost);

void orc_wall_law(const aither_cfg *cfg, int mode, int tag, const double *state,
                  double wallDist, const double area[3], int isLower, double out[11]);

void orc_mixture_transport(const aither_cfg *cfg, const double *state, double out[2]);

#ifdef __cplusplus
}
#endif
#endif
