/* ref_harness.cpp -- drives the UNMODIFIED reference solver (mnucci32/aither,
 * compiled from /root/reference by oracle/Makefile) one phase at a time and
 * dumps its arrays at full fp64 precision.
 *
 * TEST INFRASTRUCTURE ONLY: nothing under aither_b200/ links or executes this.
 * It exists because the reference's own log (.resid) prints 4 significant
 * digits (reference src/output.cpp:1050-1081), which cannot pin a 1e-12 parity
 * claim. The set-up sequence follows reference src/main.cpp:101-225 and the
 * per-iteration sequence follows mgSolution::Iterate / ImplicitUpdate
 * (reference src/mgSolution.cpp:209-269), calling the reference's own public
 * methods in the same order so each phase boundary can be observed.
 *
 * usage: aither_dump <case.inp> <out.bin> [--iters N] [--full a,b,c]
 *                    [--geom] [--time]
 *   --iters N     run N time steps (default: the .inp `iterations`)
 *   --full list   iterations (0-based) whose per-phase arrays are dumped
 *   --geom        dump grid metrics / initial state / config (block set-up)
 *   --time        print per-iteration wall time of the hot path to stdout
 *
 * Dump format ("ADMP1"): repeated records
 *   u32 nameLen | name | u8 dtype('d'|'i') | u32 ndim | i64 dims[ndim] | data
 */
#include <algorithm>
#include <array>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>
#include <random>

// The reference keeps its fields private; the harness reads them directly and is
// therefore compiled with -fno-access-control (oracle/Makefile).
#include "mpi.h"
#include "boundaryConditions.hpp"
#include "eos.hpp"
#include "fluid.hpp"
#include "gridLevel.hpp"
#include "input.hpp"
#include "inputStates.hpp"
#include "kdtree.hpp"
#include "linearSolver.hpp"
#include "logFileManager.hpp"
#include "macros.hpp"
#include "matMultiArray3d.hpp"
#include "mgSolution.hpp"
#include "multiArray3d.hpp"
#include "output.hpp"
#include "parallel.hpp"
#include "physicsModels.hpp"
#include "plot3d.hpp"
#include "procBlock.hpp"
#include "resid.hpp"
#include "thermodynamic.hpp"
#include "transport.hpp"
#include "turbulence.hpp"
#include "diffusion.hpp"
#include "utility.hpp"
#include "varArray.hpp"
#include "vector3d.hpp"

namespace {

struct Dump {
  FILE *f = nullptr;
  explicit Dump(const std::string &path) {
    f = std::fopen(path.c_str(), "wb");
    if (!f) { std::perror("open dump"); std::exit(2); }
    std::fwrite("ADMP1", 1, 5, f);
  }
  ~Dump() { if (f) std::fclose(f); }
  void header(const std::string &name, char dtype,
              const std::vector<int64_t> &dims) {
    const uint32_t nl = name.size();
    std::fwrite(&nl, 4, 1, f);
    std::fwrite(name.data(), 1, nl, f);
    std::fwrite(&dtype, 1, 1, f);
    const uint32_t nd = dims.size();
    std::fwrite(&nd, 4, 1, f);
    std::fwrite(dims.data(), 8, nd, f);
  }
  void doubles(const std::string &name, const double *p,
               const std::vector<int64_t> &dims) {
    header(name, 'd', dims);
    int64_t n = 1;
    for (auto d : dims) n *= d;
    std::fwrite(p, 8, n, f);
  }
  void ints(const std::string &name, const int *p,
            const std::vector<int64_t> &dims) {
    header(name, 'i', dims);
    int64_t n = 1;
    for (auto d : dims) n *= d;
    std::fwrite(p, 4, n, f);
  }
  void scalar(const std::string &name, double v) { doubles(name, &v, {1}); }
  void iscalar(const std::string &name, int v) { ints(name, &v, {1}); }
  void vec(const std::string &name, const std::vector<double> &v) {
    doubles(name, v.data(), {static_cast<int64_t>(v.size())});
  }
  void ivec(const std::string &name, const std::vector<int> &v) {
    ints(name, v.data(), {static_cast<int64_t>(v.size())});
  }
  // any multiArray3d whose element type is a bag of doubles
  template <typename T>
  void field(const std::string &name, const multiArray3d<T> &a) {
    static_assert(sizeof(T) % sizeof(double) == 0, "element must be doubles");
    const int64_t per = sizeof(T) / sizeof(double) * a.BlockSize();
    doubles(name, reinterpret_cast<const double *>(a.data_.data()),
            {a.NumK(), a.NumJ(), a.NumI(), per});
  }
};

int BcTypeId(const std::string &n) {
  static const std::map<std::string, int> ids = {
      {"slipWall", 1},         {"viscousWall", 2},      {"characteristic", 3},
      {"inlet", 4},            {"supersonicInflow", 5}, {"supersonicOutflow", 6},
      {"stagnationInlet", 7},  {"pressureOutlet", 8},   {"interblock", 9},
      {"periodic", 10}};
  const auto it = ids.find(n);
  return it == ids.end() ? 0 : it->second;
}

std::vector<int> ParseList(const std::string &s) {
  std::vector<int> out;
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ',')) {
    if (!tok.empty()) out.push_back(std::stoi(tok));
  }
  return out;
}

void DumpConfig(Dump &d, const input &inp, const physics &phys) {
  const int ns = inp.NumSpecies();
  d.iscalar("cfg/numEquations", inp.NumEquations());
  d.iscalar("cfg/numSpecies", ns);
  d.iscalar("cfg/numTurb", inp.NumTurbEquations());
  d.iscalar("cfg/numGhosts", inp.NumberGhostLayers());
  d.iscalar("cfg/isViscous", inp.IsViscous());
  d.iscalar("cfg/isRANS", inp.IsRANS());
  d.iscalar("cfg/isBlockMatrix", inp.IsBlockMatrix());
  d.iscalar("cfg/isImplicit", inp.IsImplicit());
  d.iscalar("cfg/isMultilevelTime", inp.IsMultilevelInTime());
  d.iscalar("cfg/matrixRequiresInit", inp.MatrixRequiresInitialization());
  d.iscalar("cfg/matrixSweeps", inp.MatrixSweeps());
  d.iscalar("cfg/nonlinearIterations", inp.NonlinearIterations());
  const auto fr = inp.FaceReconstruction();
  int recon = 1;
  if (inp.UsingConstantReconstruction()) recon = 0;
  else if (fr == "weno") recon = 2;
  else if (fr == "wenoZ") recon = 3;
  d.iscalar("cfg/recon", recon);
  const auto lim = inp.Limiter();
  d.iscalar("cfg/limiter", lim == "none" ? 0 : (lim == "vanAlbada" ? 1 : 2));
  d.iscalar("cfg/invFlux", inp.InviscidFlux() == "roe" ? 0 : 1);
  d.iscalar("cfg/invFluxJac", inp.InvFluxJac() == "rusanov" ? 0 : 1);
  d.iscalar("cfg/viscRecon",
            inp.ViscousFaceReconstruction() == "centralFourth" ? 1 : 0);
  const auto tm = inp.TurbulenceModel();
  d.iscalar("cfg/turbModel",
            tm == "none" ? 0 : (tm == "kOmegaWilcox2006" ? 1
                                : (tm == "sst2003" ? 2 : 99)));
  const auto ms = inp.MatrixSolver();
  d.iscalar("cfg/solver", (ms == "lusgs" || ms == "blusgs") ? 0 : 1);
  d.scalar("cfg/kappa", inp.Kappa());
  d.scalar("cfg/theta", inp.Theta());
  d.scalar("cfg/zeta", inp.Zeta());
  d.scalar("cfg/matrixRelaxation", inp.MatrixRelaxation());
  d.scalar("cfg/dualTimeCFL", inp.DualTimeCFL());
  d.scalar("cfg/dtNondim", inp.Dt() > 0.0 ? inp.Dt() * inp.ARef() / inp.LRef()
                                         : -1.0);
  d.scalar("cfg/viscousCFLCoeff", inp.ViscousCFLCoefficient());
  d.scalar("cfg/cflStart", inp.CFLStart());
  d.scalar("cfg/cflStep", inp.CFLStep());
  d.scalar("cfg/cflMax", inp.CFLMax());
  d.scalar("cfg/rRef", inp.RRef());
  d.scalar("cfg/tRef", inp.TRef());
  d.scalar("cfg/lRef", inp.LRef());
  d.scalar("cfg/aRef", inp.ARef());
  std::vector<double> R(ns), n(ns), hf(ns);
  for (int s = 0; s < ns; ++s) {
    R[s] = phys.Thermodynamic()->R(s);
    n[s] = phys.Thermodynamic()->N(s);
    hf[s] = phys.Thermodynamic()->Hf(s);
  }
  d.vec("cfg/gasConstant", R);
  d.vec("cfg/gasConstantEos", phys.EoS()->GasConstants());
  d.vec("cfg/n", n);
  d.vec("cfg/hf", hf);
  d.vec("cfg/mixtureRef", inp.MixtureRef());
  d.scalar("cfg/nondimScaling", phys.Transport()->NondimScaling());
  d.scalar("cfg/muMixRef", phys.Transport()->MuRef());
  // transport (sutherland) coefficients, per species
  {
    std::vector<double> vc1(ns), vs(ns), kc1(ns), ks(ns), mm(ns);
    for (int s = 0; s < ns; ++s) {
      const auto &fl = inp.Fluid(s);
      vc1[s] = fl.ViscosityCoeffs()[0];
      vs[s] = fl.ViscosityCoeffs()[1];
      kc1[s] = fl.ConductivityCoeffs()[0];
      ks[s] = fl.ConductivityCoeffs()[1];
      mm[s] = fl.MolarMass();
    }
    d.vec("cfg/suthViscC1", vc1);
    d.vec("cfg/suthViscS", vs);
    d.vec("cfg/suthCondC1", kc1);
    d.vec("cfg/suthCondS", ks);
    d.vec("cfg/molarMass", mm);
  }
  // diffusionModel: none -> no species diffusion (reference src/input.cpp:810-823)
  d.scalar("cfg/schmidt", inp.DiffusionModel() == "schmidt" ? inp.SchmidtNumber() : -1.0);
  // boundary-condition state table (already nondimensional)
  const int nb = inp.bcStates_.size();
  d.iscalar("cfg/numBCStates", nb);
  for (int b = 0; b < nb; ++b) {
    const auto &st = inp.bcStates_[b];
    const std::string p = "cfg/bc" + std::to_string(b) + "/";
    d.iscalar(p + "tag", st->Tag());
    d.iscalar(p + "endTag", st->EndTag());
    d.iscalar(p + "type", BcTypeId(st->Name()));
    d.scalar(p + "density", st->Density());
    d.scalar(p + "pressure", st->Pressure());
    const auto v = st->Velocity();
    d.vec(p + "velocity", {v.X(), v.Y(), v.Z()});
    const auto dir = st->Direction();
    d.vec(p + "direction", {dir.X(), dir.Y(), dir.Z()});
    d.scalar(p + "stagnationPressure", st->StagnationPressure());
    d.scalar(p + "stagnationTemperature", st->StagnationTemperature());
    d.scalar(p + "temperature", st->Temperature());
    d.scalar(p + "heatFlux", st->HeatFlux());
    d.iscalar(p + "isIsothermal", st->IsIsothermal());
    d.iscalar(p + "isAdiabatic", st->IsAdiabatic());
    d.iscalar(p + "isConstantHeatFlux", st->IsConstantHeatFlux());
    d.iscalar(p + "isWallLaw", st->IsWallLaw());
    d.scalar(p + "vonKarmen", st->IsWallLaw() ? st->VonKarmen() : 0.41);
    d.scalar(p + "wallConstant", st->IsWallLaw() ? st->WallConstant() : 5.5);
    d.iscalar(p + "isNonreflecting", st->IsNonreflecting());
    d.scalar(p + "lengthScale", st->LengthScale());
    d.scalar(p + "turbulenceIntensity", st->TurbulenceIntensity());
    d.scalar(p + "eddyViscosityRatio", st->EddyViscosityRatio());
    std::vector<double> mf(ns, 0.0);
    for (const auto &kv : st->MassFractions()) {
      if (inp.HaveSpecies(kv.first)) mf[inp.SpeciesIndex(kv.first)] = kv.second;
    }
    d.vec(p + "massFractions", mf);
  }
}

void DumpBlockSetup(Dump &d, const gridLevel &lvl, const std::string &pre = "") {
  d.iscalar(pre + "numBlocks", lvl.NumBlocks());
  for (int bb = 0; bb < lvl.NumBlocks(); ++bb) {
    const auto &blk = lvl.Block(bb);
    const std::string p = pre + "b" + std::to_string(bb) + "/";
    d.ivec(p + "dims", {blk.NumI(), blk.NumJ(), blk.NumK(), blk.NumGhosts(),
                        blk.ParentBlock(), blk.Rank(), blk.LocalPosition(),
                        blk.GlobalPos()});
    const auto &bc = blk.bc_;
    std::vector<int> surf;
    for (int s = 0; s < bc.NumSurfaces(); ++s) {
      surf.push_back(BcTypeId(bc.GetBCTypes(s)));
      surf.push_back(bc.GetIMin(s));
      surf.push_back(bc.GetIMax(s));
      surf.push_back(bc.GetJMin(s));
      surf.push_back(bc.GetJMax(s));
      surf.push_back(bc.GetKMin(s));
      surf.push_back(bc.GetKMax(s));
      surf.push_back(bc.GetTag(s));
      surf.push_back(bc.GetSurfaceType(s));
    }
    d.ints(p + "surfaces", surf.data(), {bc.NumSurfaces(), 9});
    d.field(p + "vol", blk.vol_);
    d.field(p + "fAreaI", blk.fAreaI_);
    d.field(p + "fAreaJ", blk.fAreaJ_);
    d.field(p + "fAreaK", blk.fAreaK_);
    d.field(p + "center", blk.center_);
    d.field(p + "fCenterI", blk.fCenterI_);
    d.field(p + "fCenterJ", blk.fCenterJ_);
    d.field(p + "fCenterK", blk.fCenterK_);
    d.field(p + "cellWidthI", blk.cellWidthI_);
    d.field(p + "cellWidthJ", blk.cellWidthJ_);
    d.field(p + "cellWidthK", blk.cellWidthK_);
    d.field(p + "wallDist", blk.wallDist_);
    d.field(p + "state0", blk.state_);
    // node coordinates of the block (no ghosts), for geometry cross-checks
    d.field(p + "nodes", blk.nodes_.coords_);
  }
  const auto &conns = lvl.Connections();
  std::vector<int> c;
  for (const auto &cn : conns) {
    c.insert(c.end(), {cn.rank_[0], cn.rank_[1], cn.block_[0], cn.block_[1],
                       cn.localBlock_[0], cn.localBlock_[1], cn.boundary_[0],
                       cn.boundary_[1], cn.d1Start_[0], cn.d1Start_[1],
                       cn.d1End_[0], cn.d1End_[1], cn.d2Start_[0],
                       cn.d2Start_[1], cn.d2End_[0], cn.d2End_[1],
                       cn.constSurf_[0], cn.constSurf_[1]});
    for (int q = 0; q < 8; ++q) c.push_back(cn.patchBorder_[q] ? 1 : 0);
    c.push_back(cn.orientation_);
    c.push_back(cn.isInterblock_ ? 1 : 0);
  }
  d.ints(pre + "connections", c.data(), {static_cast<int64_t>(conns.size()), 28});
}

// multigrid transfer maps between level `fl` and the next coarser one: fine cell -> coarse cell,
// volume weight of the fine cell in it (both kept by the fine level) and the seven trilinear
// coefficients of the fine cell centre in the coarse cell (kept by the coarse level)
void DumpTransfer(Dump &d, const gridLevel &fine, const gridLevel &coarse, int fl) {
  for (int bb = 0; bb < fine.NumBlocks(); ++bb) {
    const std::string p = "L" + std::to_string(fl) + "/b" + std::to_string(bb) + "/";
    const auto &tc = fine.toCoarse_[bb];
    std::vector<int> flat;
    for (const auto &v : tc.data_) flat.insert(flat.end(), {v.X(), v.Y(), v.Z()});
    d.ints(p + "toCoarse", flat.data(), {tc.NumK(), tc.NumJ(), tc.NumI(), 3});
    d.field(p + "volWeightFactor", fine.volWeightFactor_[bb]);
    const auto &pc = coarse.prolongCoeffs_[bb];
    std::vector<double> co;
    for (const auto &a : pc.data_) co.insert(co.end(), a.begin(), a.end());
    d.doubles(p + "prolongCoeffs", co.data(), {pc.NumK(), pc.NumJ(), pc.NumI(), 7});
  }
}

void DumpStates(Dump &d, const gridLevel &lvl, const std::string &tag) {
  for (int bb = 0; bb < lvl.NumBlocks(); ++bb) {
    d.field("b" + std::to_string(bb) + "/state@" + tag, lvl.Block(bb).state_);
  }
}

}  // namespace

int main(int argc, char *argv[]) {
  if (argc < 3) {
    std::fprintf(stderr,
                 "usage: %s case.inp out.bin [--iters N] [--full a,b] "
                 "[--geom] [--time]\n", argv[0]);
    return 2;
  }
  const std::string inputFile = argv[1];
  const std::string outFile = argv[2];
  int iters = -1;
  std::vector<int> full;
  bool geom = false, timing = false;
  for (int a = 3; a < argc; ++a) {
    const std::string s = argv[a];
    if (s == "--iters" && a + 1 < argc) iters = std::stoi(argv[++a]);
    else if (s == "--full" && a + 1 < argc) full = ParseList(argv[++a]);
    else if (s == "--geom") geom = true;
    else if (s == "--time") timing = true;
  }

  MPI_Init(&argc, &argv);
  const int rank = 0, numProcs = 1;

  // ---- set-up: same call sequence as reference main.cpp:101-225 ----------
  auto totalCells = 0.0;
  input inp(inputFile, "none");
  decomposition decomp;
  auto numProcBlock = 0;
  inp.ReadInput(rank);
  inp.NondimensionalizeFluid();
  const auto phys = inp.AssignPhysicsModels();
  inp.NondimensionalizeStateData(phys.EoS());

  residual l2First(inp.NumEquations(), inp.NumSpecies());
  mgSolution solution;
  auto mesh = ReadP3dGrid(inp.GridName(), inp.LRef(), totalCells);
  auto bcs = inp.AllBC();
  if (inp.DecompMethod() == "manual") {
    decomp = ManualDecomposition(mesh, bcs, numProcs);
  } else {
    decomp = CubicDecomposition(mesh, bcs, numProcs);
  }
  solution.ConstructFinestLevel(mesh, bcs, decomp, phys, "none", inp, l2First);
  auto viscFaces = GetViscousFaceCenters(solution.Finest().Blocks());

  MPI_Datatype MPI_vec3d, MPI_procBlockInts, MPI_connection, MPI_DOUBLE_5INT,
      MPI_vec3dMag, MPI_uncoupledScalar, MPI_tensorDouble;
  SetDataTypesMPI(MPI_vec3d, MPI_procBlockInts, MPI_connection, MPI_DOUBLE_5INT,
                  MPI_vec3dMag, MPI_uncoupledScalar, MPI_tensorDouble);
  decomp.Broadcast();
  SendNumProcBlocks(decomp.NumBlocksOnAllProc(), numProcBlock);
  auto local = solution.SendFinestGridLevel(rank, numProcBlock, MPI_vec3d,
                                            MPI_vec3dMag, MPI_connection, inp);
  local.ConstructMultigrids(decomp, inp, phys, rank, MPI_connection, MPI_vec3d,
                            MPI_vec3dMag);
  local.AuxillaryAndWidths(phys);
  kdtree tree(viscFaces);
  if (tree.Size() > 0) {
    local.CalcWallDistance(tree);
    local.SwapWallDist(rank, inp.NumberGhostLayers());
  }

  if (iters < 0) iters = inp.Iterations();
  Dump d(outFile);
  d.scalar("totalCells", totalCells);
  d.iscalar("iterations", iters);
  DumpConfig(d, inp, phys);
  auto &lvl = local[local.FinestIndex()];
  if (geom) DumpBlockSetup(d, lvl);
  d.iscalar("cfg/multigridLevels", local.NumGridLevels());
  d.iscalar("cfg/mgCycleIndex", inp.MultigridCycleIndex());
  if (geom) {  // coarse levels (prefix L<level>/) and the transfer maps between levels
    for (int ll = 1; ll < local.NumGridLevels(); ++ll) {
      DumpBlockSetup(d, local[ll], "L" + std::to_string(ll) + "/");
      DumpTransfer(d, local[ll - 1], local[ll], ll - 1);
    }
  }

  const int neq = inp.NumEquations();
  std::vector<double> histL2, histLinf, histMat, histCfl, histTime;
  std::vector<int> histLinfLoc;

  // ---- iteration loop: reference main.cpp:231-302 ------------------------
  for (int nn = 0; nn < iters; ++nn) {
    inp.CalcCFL(nn);
    local.StoreOldSolution(inp, phys, nn);
    for (int mm = 0; mm < inp.NonlinearIterations(); ++mm) {
      residual residL2(neq, inp.NumSpecies());
      resid residLinf;
      const bool dumpAll =
          mm == 0 && std::find(full.begin(), full.end(), nn) != full.end();
      const std::string it = "it" + std::to_string(nn);
      const auto t0 = std::chrono::high_resolution_clock::now();
      double matrixResid = 0.0;
      if (!dumpAll) {
        matrixResid = local.Iterate(inp, phys, MPI_tensorDouble, MPI_vec3d, mm,
                                    rank, residL2, residLinf);
      } else {
        // same order as mgSolution::Iterate + ImplicitUpdate + CycleAtLevel
        // for a single grid level, observed after every phase
        d.scalar(it + "/cfl", inp.CFL());
        DumpStates(d, lvl, it + ".start");
        lvl.GetBoundaryConditions(inp, phys, rank);
        DumpStates(d, lvl, it + ".bc");
        lvl.CalcResidual(phys, inp, rank, MPI_tensorDouble, MPI_vec3d);
        for (int bb = 0; bb < lvl.NumBlocks(); ++bb) {
          const auto p = "b" + std::to_string(bb) + "/";
          d.field(p + "residual@" + it, lvl.Block(bb).residual_);
          d.field(p + "specRadius@" + it, lvl.Block(bb).specRadius_);
          d.field(p + "temperature@" + it, lvl.Block(bb).temperature_);
          if (inp.IsViscous()) {
            d.field(p + "viscosity@" + it, lvl.Block(bb).viscosity_);
            d.field(p + "velocityGrad@" + it, lvl.Block(bb).velocityGrad_);
            d.field(p + "state@" + it + ".viscbc", lvl.Block(bb).state_);
          }
          if (inp.IsTurbulent()) {
            d.field(p + "eddyViscosity@" + it, lvl.Block(bb).eddyViscosity_);
          }
          if (inp.IsRANS()) {
            d.field(p + "f1@" + it, lvl.Block(bb).f1_);
            d.field(p + "f2@" + it, lvl.Block(bb).f2_);
            d.field(p + "tkeGrad@" + it, lvl.Block(bb).tkeGrad_);
            d.field(p + "omegaGrad@" + it, lvl.Block(bb).omegaGrad_);
          }
          d.field(p + "diagRaw@" + it, lvl.solver_->a_[bb]);
          // wall variables of every viscous-wall surface (include/wallData.hpp:40-57), as the
          // viscous fluxes of this evaluation left them: 12 doubles per wall face
          for (int ww = 0; ww < static_cast<int>(lvl.Block(bb).wallData_.size()); ++ww) {
            const auto &wd = lvl.Block(bb).wallData_[ww];
            const auto &sf = wd.surf_;
            d.ivec(p + "wall" + std::to_string(ww) + "/surface",
                   {sf.IMin(), sf.IMax(), sf.JMin(), sf.JMax(), sf.KMin(), sf.KMax(),
                    sf.SurfaceType(), sf.Tag()});
            std::vector<double> v;
            for (const auto &w : wd.data_.data_) {
              v.insert(v.end(), {w.yplus_, w.shearStress_.X(), w.shearStress_.Y(),
                                 w.shearStress_.Z(), w.heatFlux_, w.temperature_,
                                 w.turbEddyVisc_, w.viscosity_, w.density_,
                                 w.frictionVelocity_, w.tke_, w.sdr_});
            }
            d.doubles(p + "wall" + std::to_string(ww) + "/vars@" + it, v.data(),
                      {wd.NumK(), wd.NumJ(), wd.NumI(), 12});
          }
        }
        lvl.CalcTimeStep(inp);
        for (int bb = 0; bb < lvl.NumBlocks(); ++bb) {
          d.field("b" + std::to_string(bb) + "/dt@" + it, lvl.Block(bb).dt_);
        }
        if (inp.IsImplicit()) {
          lvl.InvertDiagonal(inp);
          lvl.InitializeMatrixUpdate(inp, phys);
          for (int bb = 0; bb < lvl.NumBlocks(); ++bb) {
            const auto p = "b" + std::to_string(bb) + "/";
            d.field(p + "diag@" + it, lvl.solver_->a_[bb]);
            d.field(p + "diagInv@" + it, lvl.solver_->aInv_[bb]);
            d.field(p + "x0@" + it, lvl.solver_->x_[bb]);
          }
          auto mr = lvl.Relax(phys, inp, rank, inp.MatrixSweeps());
          auto l2 = 0.0;
          auto totalSize = 0;
          for (int bb = 0; bb < lvl.NumBlocks(); ++bb) {
            const auto p = "b" + std::to_string(bb) + "/";
            d.field(p + "x@" + it, lvl.solver_->x_[bb]);
            d.field(p + "matrixResid@" + it, mr[bb]);
          }
          for (auto &m : mr) {
            m *= m;
            l2 += std::accumulate(std::begin(m), std::end(m), 0.0);
            totalSize += m.Size();
          }
          matrixResid = l2 / totalSize;
          lvl.UpdateBlocks(inp, phys, mm, residL2, residLinf);
          lvl.ResetDiagonal();
        } else {
          lvl.ExplicitUpdate(inp, phys, mm, residL2, residLinf);
        }
        DumpStates(d, lvl, it + ".end");
      }
      const auto t1 = std::chrono::high_resolution_clock::now();
      const double sec = std::chrono::duration<double>(t1 - t0).count();
      for (int e = 0; e < neq; ++e) histL2.push_back(residL2[e]);
      histLinf.push_back(residLinf.Linf());
      histLinfLoc.insert(histLinfLoc.end(),
                         {residLinf.Block(), residLinf.ILoc(), residLinf.JLoc(),
                          residLinf.KLoc(), residLinf.Eqn()});
      histMat.push_back(matrixResid);
      histCfl.push_back(inp.CFL());
      histTime.push_back(sec);
      if (timing) {
        std::printf("iter %d nl %d time_s %.6f\n", nn, mm, sec);
        std::fflush(stdout);
      }
    }
  }
  const int64_t nrec = histMat.size();
  d.doubles("hist/residL2", histL2.data(), {nrec, neq});
  d.doubles("hist/linf", histLinf.data(), {nrec});
  d.ints("hist/linfLoc", histLinfLoc.data(), {nrec, 5});
  d.doubles("hist/matrixResid", histMat.data(), {nrec});
  d.doubles("hist/cfl", histCfl.data(), {nrec});
  d.doubles("hist/time", histTime.data(), {nrec});
  DumpStates(d, lvl, "final");
  return 0;
}
