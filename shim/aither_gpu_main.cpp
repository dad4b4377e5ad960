/* aither_gpu_main.cpp -- the reference's main program with the iteration body on the B200.
 *
 * Everything outside the marked lines is the call sequence of the reference's own main()
 * (src/main.cpp:56-310: input, physics, grid, decomposition, multigrid levels, wall distance,
 * residual log, output) made by the reference's own, unmodified objects (oracle/_ref/obj/*.o).
 * The three places a maintainer changes are marked  // <<< gpuPath:
 *   1. construct the device path once the levels exist
 *   2. StoreOldSolution / Iterate go to it
 *   3. the state comes back when output or a restart file is written
 * Built by `make -C oracle shim` into oracle/_ref/aither_gpu_main; run by tests/test_gpu_shim.py on
 * the shipped .inp files (the last line of the .resid file against the reference's regression
 * goldens, testCases/regressionTests.py).
 *
 * usage: aither_gpu_main case.inp     (single rank: the container has no MPI -- oracle/stub/mpi.h;
 *        with MPI every rank builds its own gpuPath on its own GPU and passes an ncclComm_t) */
#include <cmath>
#include <iostream>
#include <string>
#include <vector>

#include "mpi.h"
#include "boundaryConditions.hpp"
#include "gridLevel.hpp"
#include "input.hpp"
#include <cstdlib>

#include "kdtree.hpp"
#include "logFileManager.hpp"
#include "macros.hpp"
#include "mgSolution.hpp"
#include "output.hpp"
#include "parallel.hpp"
#include "physicsModels.hpp"
#include "plot3d.hpp"
#include "procBlock.hpp"
#include "resid.hpp"
#include "utility.hpp"
#include "varArray.hpp"

#include "gpuPath.hpp"

using std::cout;
using std::cerr;
using std::endl;
using std::string;
using std::vector;

int main(int argc, char *argv[]) {
  auto rank = 0;
  auto numProcs = 1;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &numProcs);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  if (argc != 2) {
    cerr << "USAGE: aither_gpu_main inputFile.inp" << endl;
    return EXIT_FAILURE;
  }
  string inputFile = argv[1];
  string restartFile = "none";

  auto totalCells = 0.0;
  input inp(inputFile, restartFile);
  decomposition decomp;
  auto numProcBlock = 0;
  inp.ReadInput(rank);
  logFileManager logs(inp, rank);
  inp.NondimensionalizeFluid();
  const auto phys = inp.AssignPhysicsModels();
  inp.NondimensionalizeStateData(phys.EoS());

  mgSolution solution;  // only keep finest grid level globally
  vector<vector3d<double>> viscFaces;
  if (rank == ROOTP) {
    auto mesh = ReadP3dGrid(inp.GridName(), inp.LRef(), totalCells);
    auto bcs = inp.AllBC();
    if (inp.DecompMethod() == "manual") {
      decomp = ManualDecomposition(mesh, bcs, numProcs);
    } else {
      decomp = CubicDecomposition(mesh, bcs, numProcs);
    }
    solution.ConstructFinestLevel(mesh, bcs, decomp, phys, restartFile, inp, logs.L2First());
    viscFaces = GetViscousFaceCenters(solution.Finest().Blocks());
  }
  MPI_Datatype MPI_vec3d, MPI_procBlockInts, MPI_connection, MPI_DOUBLE_5INT, MPI_vec3dMag,
      MPI_uncoupledScalar, MPI_tensorDouble;
  SetDataTypesMPI(MPI_vec3d, MPI_procBlockInts, MPI_connection, MPI_DOUBLE_5INT, MPI_vec3dMag,
                  MPI_uncoupledScalar, MPI_tensorDouble);
  decomp.Broadcast();
  SendNumProcBlocks(decomp.NumBlocksOnAllProc(), numProcBlock);
  auto localSolution = solution.SendFinestGridLevel(rank, numProcBlock, MPI_vec3d, MPI_vec3dMag,
                                                    MPI_connection, inp);
  localSolution.ConstructMultigrids(decomp, inp, phys, rank, MPI_connection, MPI_vec3d,
                                    MPI_vec3dMag);
  localSolution.AuxillaryAndWidths(phys);
  BroadcastViscFaces(MPI_vec3d, viscFaces);
  // AITHER_GPU_WALLDIST set: the wall distance is computed on the device (gpuPath 1b below)
  // instead of by the k-d tree search here
  const bool gpuWallDist = std::getenv("AITHER_GPU_WALLDIST") != nullptr;
  if (!gpuWallDist) {
    kdtree tree(viscFaces);
    if (tree.Size() > 0) {
      localSolution.CalcWallDistance(tree);
      localSolution.SwapWallDist(rank, inp.NumberGhostLayers());
    }
  }
  solution.GetFinestGridLevel(localSolution, rank, MPI_uncoupledScalar, MPI_vec3d,
                              MPI_tensorDouble, inp);
  if (rank == ROOTP) {
    WriteCellCenter(inp.GridName(), solution.Finest().Blocks(), decomp, inp);
    WriteOutput(solution.Finest().Blocks(), phys, inp.IterationStart(), decomp, inp);
  }

  // <<< gpuPath 1: every grid level of this rank on the device
  gpuPath gpu(inp, phys, localSolution, rank, numProcs);
  if (gpuWallDist) {  // <<< gpuPath 1b (was CalcWallDistance + SwapWallDist above)
    gpu.ComputeWallDistance(viscFaces);
    cout << "gpuPath: wall distance from " << viscFaces.size() << " wall faces on the device" << endl;
  }
  cout << "gpuPath: " << aither_gpu_version() << ", " << localSolution.NumGridLevels()
       << " grid level(s), " << localSolution[0].NumBlocks() << " block(s)" << endl;

  for (auto nn = 0; nn < inp.Iterations(); ++nn) {
    logs.GetIterStart();
    inp.CalcCFL(nn);
    gpu.StoreOldSolution(nn);  // <<< gpuPath 2 (was localSolution.StoreOldSolution)
    for (auto mm = 0; mm < inp.NonlinearIterations(); ++mm) {
      residual residL2(inp.NumEquations(), inp.NumSpecies());
      resid residLinf;
      // <<< gpuPath 2 (was localSolution.Iterate)
      auto matrixResid = gpu.Iterate(inp, mm, residL2, residLinf);
      // (multi-rank: residL2.GlobalReduceMPI, residLinf.GlobalReduceMPI, MPI_Reduce as in main.cpp)
      if (rank == ROOTP) {
        residL2.SquareRoot();
        matrixResid = sqrt(matrixResid / (totalCells * inp.NumEquations()));
        logs.WriteResiduals(inp, residL2, residLinf, matrixResid, nn + inp.IterationStart(), mm);
      }
    }
    if (inp.WriteOutput(nn) || inp.WriteRestart(nn)) {
      gpu.DownloadStates(localSolution);  // <<< gpuPath 3
      solution.GetFinestGridLevel(localSolution, rank, MPI_uncoupledScalar, MPI_vec3d,
                                  MPI_tensorDouble, inp);
      if (rank == ROOTP && inp.WriteOutput(nn)) {
        WriteOutput(solution.Finest().Blocks(), phys, (nn + inp.IterationStart() + 1), decomp, inp);
      }
      if (rank == ROOTP && inp.WriteRestart(nn)) {
        WriteRestart(solution.Finest().Blocks(), phys, (nn + inp.IterationStart() + 1), decomp, inp,
                     logs.L2First());
      }
    }
    logs.WriteTime(nn);
  }
  if (rank == ROOTP) cout << endl << "Program Complete" << endl;
  MPI_Finalize();
  return EXIT_SUCCESS;
}
