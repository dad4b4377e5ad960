/* gpuPath.hpp -- the reference-side shim: mgSolution's per-iteration work done by the B200
 * library (include/aither_gpu.h) behind the reference's own objects.
 *
 * This is the file a maintainer adds to mnucci32/aither (v0.10.0). It is compiled HERE against
 * the unmodified reference objects (oracle/Makefile: `make shim` -> oracle/_ref/aither_gpu_main)
 * and run by tests/test_gpu_shim.py on the shipped .inp files. What it replaces:
 *
 *   gpuPath::gpuPath            gridLevel / linearSolver construction for every grid level
 *                               (src/gridLevel.cpp:50-120, src/input.cpp:843-858): POD snapshots
 *                               of input + physics (aither_cfg), of every procBlock
 *                               (aither_block_desc: host arrays in the reference's own layout)
 *                               and of every connection (aither_conn), multigrid transfer maps
 *   gpuPath::StoreOldSolution   mgSolution::StoreOldSolution        src/mgSolution.cpp:103-114
 *   gpuPath::Iterate            mgSolution::Iterate / ImplicitUpdate / CycleAtLevel
 *                                                                   src/mgSolution.cpp:160-269
 *   gpuPath::ComputeWallDistance mgSolution::CalcWallDistance + SwapWallDist (optional)
 *                                                                   src/main.cpp:191-202
 *   gpuPath::DownloadStates     the state back into procBlock::state_ when main.cpp writes
 *                               output or a restart file            src/main.cpp:280-300
 *
 * procBlock / gridLevel / connection keep their arrays private; the shim reads them directly and
 * is therefore compiled with -fno-access-control (in the reference tree: `friend class gpuPath;`
 * in procBlock, gridLevel, connection, boundaryConditions and input).
 */
#ifndef AITHER_GPU_PATH_HPP
#define AITHER_GPU_PATH_HPP

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

#include "aither_gpu.h"
#include "boundaryConditions.hpp"
#include "gridLevel.hpp"
#include "input.hpp"
#include "inputStates.hpp"
#include "mgSolution.hpp"
#include "physicsModels.hpp"
#include "procBlock.hpp"
#include "resid.hpp"
#include "thermodynamic.hpp"
#include "transport.hpp"
#include "varArray.hpp"

class gpuPath {
  std::vector<aither_gpu *> levels_;  // finest first
  int neq_ = 0;
  int sweeps_ = 0;
  int cycleIndex_ = 1;

  static void Check(int rc, const char *what) {
    if (rc != 0) {  // the reference's error behaviour: message + exit
      std::fprintf(stderr, "ERROR: %s: %s\n", what, aither_gpu_last_error());
      std::exit(EXIT_FAILURE);
    }
  }
  static int BcTypeId(const std::string &n) {  // aither_bc_type of a BC name
    static const std::map<std::string, int> ids = {
        {"slipWall", AITHER_BC_SLIP_WALL},
        {"viscousWall", AITHER_BC_VISCOUS_WALL},
        {"characteristic", AITHER_BC_CHARACTERISTIC},
        {"inlet", AITHER_BC_INLET},
        {"supersonicInflow", AITHER_BC_SUPERSONIC_INFLOW},
        {"supersonicOutflow", AITHER_BC_SUPERSONIC_OUTFLOW},
        {"stagnationInlet", AITHER_BC_STAGNATION_INLET},
        {"pressureOutlet", AITHER_BC_PRESSURE_OUTLET},
        {"interblock", AITHER_BC_INTERBLOCK},
        {"periodic", AITHER_BC_PERIODIC}};
    const auto it = ids.find(n);
    return it == ids.end() ? 0 : it->second;
  }

  // input + physics -> aither_cfg (src/input.cpp:674-721,1110-1144; include/inputStates.hpp)
  static aither_cfg MakeCfg(const input &inp, const physics &phys) {
    aither_cfg c = {};
    const int ns = inp.NumSpecies();
    if (ns > AITHER_MAX_SPECIES) {
      std::fprintf(stderr, "ERROR: gpuPath: more species than the library is built for\n");
      std::exit(EXIT_FAILURE);
    }
    c.numSpecies = ns;
    c.numTurb = inp.NumTurbEquations();
    c.numGhosts = inp.NumberGhostLayers();
    c.isViscous = inp.IsViscous();
    c.isRANS = inp.IsRANS();
    c.isBlockMatrix = inp.IsBlockMatrix();
    c.isMultilevelTime = inp.IsMultilevelInTime();
    const auto fr = inp.FaceReconstruction();
    c.recon = inp.UsingConstantReconstruction()
                  ? AITHER_RECON_CONSTANT
                  : (fr == "weno" ? AITHER_RECON_WENO
                                  : (fr == "wenoZ" ? AITHER_RECON_WENOZ : AITHER_RECON_MUSCL));
    const auto lim = inp.Limiter();
    c.limiter = lim == "none" ? AITHER_LIMITER_NONE
                              : (lim == "vanAlbada" ? AITHER_LIMITER_VAN_ALBADA
                                                    : AITHER_LIMITER_MINMOD);
    c.invFlux = inp.InviscidFlux() == "roe" ? AITHER_FLUX_ROE : AITHER_FLUX_AUSM;
    c.invFluxJac = inp.InvFluxJac() == "rusanov" ? AITHER_JAC_RUSANOV : AITHER_JAC_APPROX_ROE;
    c.viscRecon = inp.ViscousFaceReconstruction() == "centralFourth" ? 1 : 0;
    const auto tm = inp.TurbulenceModel();
    c.turbModel = tm == "none" ? AITHER_TURB_NONE
                               : (tm == "kOmegaWilcox2006" ? AITHER_TURB_KW_WILCOX
                                                           : (tm == "sst2003" ? AITHER_TURB_SST : 99));
    const auto ms = inp.MatrixSolver();
    c.solver = (ms == "lusgs" || ms == "blusgs") ? AITHER_SOLVER_LUSGS : AITHER_SOLVER_DPLUR;
    c.matrixSweeps = inp.MatrixSweeps();
    c.matrixRequiresInit = inp.MatrixRequiresInitialization();
    c.nonlinearIterations = inp.NonlinearIterations();
    c.kappa = inp.Kappa();
    c.theta = inp.Theta();
    c.zeta = inp.Zeta();
    c.matrixRelaxation = inp.MatrixRelaxation();
    c.dualTimeCFL = inp.DualTimeCFL();
    c.dtNondim = inp.Dt() > 0.0 ? inp.Dt() * inp.ARef() / inp.LRef() : -1.0;  // src/procBlock.cpp:808
    c.viscousCFLCoeff = inp.ViscousCFLCoefficient();
    for (int s = 0; s < ns; ++s) {
      c.gasConstant[s] = phys.Thermodynamic()->R(s);
      c.n[s] = phys.Thermodynamic()->N(s);
      c.hf[s] = phys.Thermodynamic()->Hf(s);
      const auto &fl = inp.Fluid(s);
      c.suthViscC1[s] = fl.ViscosityCoeffs()[0];
      c.suthViscS[s] = fl.ViscosityCoeffs()[1];
      c.suthCondC1[s] = fl.ConductivityCoeffs()[0];
      c.suthCondS[s] = fl.ConductivityCoeffs()[1];
      c.molarMass[s] = fl.MolarMass();
    }
    c.nondimScaling = phys.Transport()->NondimScaling();
    c.tRef = inp.TRef();
    c.muMixRef = phys.Transport()->MuRef();
    // sutherland::sutherland (src/transport.cpp:65-68): kNonDim = aRef^2 muRef / tRef
    c.kMixRef = inp.ARef() * inp.ARef() * c.muMixRef / inp.TRef();
    // diffusionModel: none -> no species diffusion (src/input.cpp:810-823)
    c.schmidt = inp.DiffusionModel() == "schmidt" ? inp.SchmidtNumber() : -1.0;
    c.turbPrandtl = 0.9;
    const int nb = static_cast<int>(inp.bcStates_.size());
    if (nb > AITHER_MAX_BC_STATES) {
      std::fprintf(stderr, "ERROR: gpuPath: more boundary states than the library's table holds\n");
      std::exit(EXIT_FAILURE);
    }
    c.numBCStates = nb;
    for (int b = 0; b < nb; ++b) {
      const auto &st = inp.bcStates_[b];
      aither_bc_state &o = c.bcStates[b];
      o.tag = st->Tag();
      o.type = BcTypeId(st->Name());
      o.density = st->Density();
      o.pressure = st->Pressure();
      const auto v = st->Velocity();
      o.velocity[0] = v.X(); o.velocity[1] = v.Y(); o.velocity[2] = v.Z();
      const auto dir = st->Direction();
      o.direction[0] = dir.X(); o.direction[1] = dir.Y(); o.direction[2] = dir.Z();
      o.stagnationPressure = st->StagnationPressure();
      o.stagnationTemperature = st->StagnationTemperature();
      o.temperature = st->Temperature();
      o.heatFlux = st->HeatFlux();
      o.isIsothermal = st->IsIsothermal();
      o.isConstantHeatFlux = st->IsConstantHeatFlux();
      o.isWallLaw = st->IsWallLaw();
      o.vonKarmen = st->IsWallLaw() ? st->VonKarmen() : 0.41;
      o.wallConstant = st->IsWallLaw() ? st->WallConstant() : 5.5;
      o.isNonreflecting = st->IsNonreflecting();
      o.lengthScale = st->LengthScale();
      o.turbulenceIntensity = st->TurbulenceIntensity();
      o.eddyViscosityRatio = st->EddyViscosityRatio();
      for (const auto &kv : st->MassFractions())
        if (inp.HaveSpecies(kv.first)) o.massFractions[inp.SpeciesIndex(kv.first)] = kv.second;
    }
    return c;
  }

  template <typename T>
  static const double *Raw(const multiArray3d<T> &a) {
    static_assert(sizeof(T) % sizeof(double) == 0, "element must be a bag of doubles");
    return reinterpret_cast<const double *>(a.data_.data());
  }

  // one grid level -> device handle
  static aither_gpu *MakeLevel(const aither_cfg &cfg, const gridLevel &lvl, int rank, int numProcs,
                               void *ncclComm, int device) {
    const int nb = lvl.NumBlocks();
    std::vector<std::vector<aither_surface>> surfs(nb);
    std::vector<aither_block_desc> descs(nb);
    for (int bb = 0; bb < nb; ++bb) {
      const procBlock &blk = lvl.Block(bb);
      const auto &bc = blk.bc_;
      for (int s = 0; s < bc.NumSurfaces(); ++s) {
        aither_surface sf;
        sf.type = BcTypeId(bc.GetBCTypes(s));
        sf.imin = bc.GetIMin(s); sf.imax = bc.GetIMax(s);
        sf.jmin = bc.GetJMin(s); sf.jmax = bc.GetJMax(s);
        sf.kmin = bc.GetKMin(s); sf.kmax = bc.GetKMax(s);
        sf.tag = bc.GetTag(s);
        surfs[bb].push_back(sf);
      }
      aither_block_desc &d = descs[bb];
      d = aither_block_desc{};
      d.ni = blk.NumI(); d.nj = blk.NumJ(); d.nk = blk.NumK();
      d.parentBlock = blk.ParentBlock();
      d.globalPos = blk.GlobalPos();
      d.numSurfaces = static_cast<int>(surfs[bb].size());
      d.surfaces = surfs[bb].data();
      d.state = Raw(blk.state_);
      d.vol = Raw(blk.vol_);
      d.fAreaI = Raw(blk.fAreaI_);
      d.fAreaJ = Raw(blk.fAreaJ_);
      d.fAreaK = Raw(blk.fAreaK_);
      d.center = Raw(blk.center_);
      d.cellWidthI = Raw(blk.cellWidthI_);
      d.cellWidthJ = Raw(blk.cellWidthJ_);
      d.cellWidthK = Raw(blk.cellWidthK_);
      d.wallDist = cfg.isViscous ? Raw(blk.wallDist_) : nullptr;
    }
    std::vector<aither_conn> conns;
    for (const auto &cn : lvl.Connections()) {  // the 12 fields verbatim
      aither_conn k = {};
      for (int s = 0; s < 2; ++s) {
        k.rank[s] = cn.rank_[s];
        k.block[s] = cn.block_[s];
        k.localBlock[s] = cn.localBlock_[s];
        k.boundary[s] = cn.boundary_[s];
        k.d1Start[s] = cn.d1Start_[s];
        k.d1End[s] = cn.d1End_[s];
        k.d2Start[s] = cn.d2Start_[s];
        k.d2End[s] = cn.d2End_[s];
        k.constSurf[s] = cn.constSurf_[s];
      }
      for (int q = 0; q < 8; ++q) k.patchBorder[q] = cn.patchBorder_[q] ? 1 : 0;
      k.orientation = cn.orientation_;
      k.isInterblock = cn.isInterblock_ ? 1 : 0;
      conns.push_back(k);
    }
    aither_gpu *h = nullptr;
    Check(aither_gpu_create(&cfg, nb, descs.data(), static_cast<int>(conns.size()), conns.data(),
                            rank, numProcs, ncclComm, device, &h),
          "aither_gpu_create");
    return h;
  }

  // mgSolution::CycleAtLevel (src/mgSolution.cpp:160-207); returns sum(mr^2) / size of level fl
  double Cycle(int fl, int mm, double cfl) {
    double mr = 0.0;
    if (fl == static_cast<int>(levels_.size()) - 1) {
      Check(aither_gpu_relax(levels_[fl], sweeps_, &mr), "aither_gpu_relax");
      return mr;
    }
    const int half = std::max(sweeps_ / 2, 1);
    Check(aither_gpu_relax(levels_[fl], half, &mr), "aither_gpu_relax");
    Check(aither_gpu_mg_restrict(levels_[fl], levels_[fl + 1], mm, cfl), "aither_gpu_mg_restrict");
    Check(aither_gpu_mg_save_update(levels_[fl + 1]), "aither_gpu_mg_save_update");
    for (int ii = 0; ii < cycleIndex_; ++ii) Cycle(fl + 1, mm, cfl);
    Check(aither_gpu_mg_subtract_saved(levels_[fl + 1]), "aither_gpu_mg_subtract_saved");
    Check(aither_gpu_mg_prolong(levels_[fl + 1], levels_[fl]), "aither_gpu_mg_prolong");
    Check(aither_gpu_relax(levels_[fl], half, &mr), "aither_gpu_relax");
    return mr;
  }

 public:
  gpuPath(const input &inp, const physics &phys, const mgSolution &sol, int rank, int numProcs,
          void *ncclComm = nullptr, int device = 0) {
    if (!inp.IsImplicit()) {
      std::fprintf(stderr, "ERROR: gpuPath: the B200 path covers the implicit time integrators\n");
      std::exit(EXIT_FAILURE);
    }
    const aither_cfg cfg = MakeCfg(inp, phys);
    neq_ = inp.NumEquations();
    sweeps_ = inp.MatrixSweeps();
    cycleIndex_ = inp.MultigridCycleIndex();
    for (int ll = 0; ll < sol.NumGridLevels(); ++ll)
      levels_.push_back(MakeLevel(cfg, sol[ll], rank, numProcs, ncclComm, device));
    // transfer maps between level pairs (include/gridLevel.hpp:56-58)
    for (int ll = 0; ll + 1 < sol.NumGridLevels(); ++ll) {
      const gridLevel &fine = sol[ll], &coarse = sol[ll + 1];
      for (int bb = 0; bb < fine.NumBlocks(); ++bb) {
        std::vector<int> tc;
        for (const auto &v : fine.toCoarse_[bb].data_) tc.insert(tc.end(), {v.X(), v.Y(), v.Z()});
        std::vector<double> pc;
        for (const auto &a : coarse.prolongCoeffs_[bb].data_) pc.insert(pc.end(), a.begin(), a.end());
        Check(aither_gpu_set_transfer(levels_[ll], bb, tc.data(), Raw(fine.volWeightFactor_[bb]),
                                      pc.data()),
              "aither_gpu_set_transfer");
      }
    }
  }
  ~gpuPath() {
    for (auto *h : levels_) aither_gpu_destroy(h);
  }
  gpuPath(const gpuPath &) = delete;
  gpuPath &operator=(const gpuPath &) = delete;

  // mgSolution::StoreOldSolution
  void StoreOldSolution(int nn) {
    Check(aither_gpu_store_old_solution(levels_[0], nn), "aither_gpu_store_old_solution");
  }

  // mgSolution::Iterate: returns the matrix residual as CycleAtLevel does, accumulates the
  // un-rooted L2 sums and the L-infinity record exactly as UpdateBlocks does
  double Iterate(const input &inp, int mm, residual &residL2, resid &residLinf) {
    std::vector<double> l2(neq_, 0.0);
    aither_linf linf = {};
    double mr = 0.0;
    const double cfl = inp.CFL();
    if (levels_.size() == 1) {
      Check(aither_gpu_iterate(levels_[0], cfl, mm, l2.data(), &linf, &mr), "aither_gpu_iterate");
    } else {
      aither_gpu *f = levels_[0];
      Check(aither_gpu_get_boundary_conditions(f), "aither_gpu_get_boundary_conditions");
      Check(aither_gpu_calc_residual(f), "aither_gpu_calc_residual");
      Check(aither_gpu_calc_time_step(f, cfl), "aither_gpu_calc_time_step");
      Check(aither_gpu_invert_diagonal(f), "aither_gpu_invert_diagonal");
      Check(aither_gpu_initialize_matrix_update(f), "aither_gpu_initialize_matrix_update");
      mr = Cycle(0, mm, cfl);
      Check(aither_gpu_update_blocks(f, mm, l2.data(), &linf), "aither_gpu_update_blocks");
      for (auto *h : levels_) Check(aither_gpu_reset_diagonal(h), "aither_gpu_reset_diagonal");
    }
    for (int e = 0; e < neq_; ++e) residL2[e] += l2[e];
    if (linf.linf > residLinf.Linf())
      residLinf = resid(linf.linf, linf.block, linf.i, linf.j, linf.k, linf.eqn);
    return mr;
  }

  // wall distance of every grid level on the device: replaces mgSolution::CalcWallDistance(tree)
  // and SwapWallDist (src/main.cpp:191-202); `viscFaces` is what main.cpp broadcasts to every rank
  // (GetViscousFaceCenters, src/utility.cpp:310-368). The procBlocks' own wallDist_ arrays are
  // left as they are: nothing on the host reads them during the run.
  void ComputeWallDistance(const std::vector<vector3d<double>> &viscFaces) {
    static_assert(sizeof(vector3d<double>) == 3 * sizeof(double), "vector3d is three doubles");
    for (aither_gpu *lv : levels_)
      Check(aither_gpu_compute_wall_distance(lv, reinterpret_cast<const double *>(viscFaces.data()),
                                             static_cast<long long>(viscFaces.size())),
            "aither_gpu_compute_wall_distance");
  }

  // state of the finest level back into the procBlocks (main.cpp writes output / restart from it)
  void DownloadStates(mgSolution &sol) {
    gridLevel &lvl = sol[sol.FinestIndex()];
    for (int bb = 0; bb < lvl.NumBlocks(); ++bb)
      Check(aither_gpu_download_state(levels_[0], bb, lvl.Block(bb).state_.data_.data()),
            "aither_gpu_download_state");
  }
};

#endif
