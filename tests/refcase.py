"""Run the reference harness (oracle/_ref/aither_dump) on a case and turn dumps into Problems.

TEST INFRASTRUCTURE ONLY. The harness binary is built here by oracle/Makefile from /root/reference
and travels to the GPU box prebuilt; /root/reference itself is never read at test time.
"""
import os
import shutil
import subprocess
import tempfile

import numpy as np

from aither_b200 import ctypes_abi as abi
from aither_b200.problem import Block, Problem, make_cfg
from refdump import read_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "aither_dump")

AIR_DAT = """n: 2.5
molarMass: 28.97
vibrationalTemperature: [3056.0]
heatOfFormation: 0
referencePressure: 101325
referenceTemperature: 298.15
referenceEntropy: 0
sutherlandViscosityC1: 1.458e-6
sutherlandViscosityS: 110.4
sutherlandConductivityC1: 2.495e-3
sutherlandConductivityS: 194.0
"""


def have_harness():
    return os.path.exists(HARNESS) and os.access(HARNESS, os.X_OK)


def run_harness(case_dir, inp_name, iters, full=(), geom=True, timing=False):
    """Run the reference on `case_dir/inp_name`; returns the dump dict."""
    out = os.path.join(case_dir, "dump.bin")
    if not os.path.exists(os.path.join(case_dir, "air.dat")):
        with open(os.path.join(case_dir, "air.dat"), "w") as f:
            f.write(AIR_DAT)
    cmd = [HARNESS, inp_name, out, "--iters", str(iters)]
    if full:
        cmd += ["--full", ",".join(str(i) for i in full)]
    if geom:
        cmd.append("--geom")
    if timing:
        cmd.append("--time")
    env = dict(os.environ)
    env.setdefault("AITHER_INSTALL_DIRECTORY", case_dir)
    res = subprocess.run(cmd, cwd=case_dir, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError("reference harness failed:\n" + res.stdout[-4000:])
    d = read_dump(out)
    d["__stdout__"] = res.stdout
    return d


def cfg_from_dump(d):
    g = lambda k: d["cfg/" + k]
    ns = int(g("numSpecies")[0])
    bcs = []
    for b in range(int(g("numBCStates")[0])):
        p = "bc%d/" % b
        bcs.append(dict(
            tag=int(g(p + "tag")[0]), type=int(g(p + "type")[0]),
            density=float(g(p + "density")[0]), velocity=list(g(p + "velocity")),
            pressure=float(g(p + "pressure")[0]), massFractions=list(g(p + "massFractions")),
            stagnationPressure=float(g(p + "stagnationPressure")[0]),
            stagnationTemperature=float(g(p + "stagnationTemperature")[0]),
            direction=list(g(p + "direction")), temperature=float(g(p + "temperature")[0]),
            heatFlux=float(g(p + "heatFlux")[0]), isIsothermal=int(g(p + "isIsothermal")[0]),
            isConstantHeatFlux=int(g(p + "isConstantHeatFlux")[0]),
            turbulenceIntensity=float(g(p + "turbulenceIntensity")[0]),
            eddyViscosityRatio=float(g(p + "eddyViscosityRatio")[0]),
            isWallLaw=int(g(p + "isWallLaw")[0]) if ("cfg/" + p + "isWallLaw") in d else 0,
            vonKarmen=float(g(p + "vonKarmen")[0]) if ("cfg/" + p + "vonKarmen") in d else 0.41,
            wallConstant=float(g(p + "wallConstant")[0]) if ("cfg/" + p + "wallConstant") in d
            else 5.5,
            isNonreflecting=int(g(p + "isNonreflecting")[0])
            if ("cfg/" + p + "isNonreflecting") in d else 0,
            lengthScale=float(g(p + "lengthScale")[0]) if ("cfg/" + p + "lengthScale") in d
            else 0.0))
    return make_cfg(
        numSpecies=ns, numTurb=int(g("numTurb")[0]), numGhosts=int(g("numGhosts")[0]),
        isViscous=int(g("isViscous")[0]), isRANS=int(g("isRANS")[0]),
        isBlockMatrix=int(g("isBlockMatrix")[0]), isMultilevelTime=int(g("isMultilevelTime")[0]),
        recon=int(g("recon")[0]), limiter=int(g("limiter")[0]), invFlux=int(g("invFlux")[0]),
        invFluxJac=int(g("invFluxJac")[0]), viscRecon=int(g("viscRecon")[0]),
        turbModel=int(g("turbModel")[0]), solver=int(g("solver")[0]),
        matrixSweeps=int(g("matrixSweeps")[0]), matrixRequiresInit=int(g("matrixRequiresInit")[0]),
        nonlinearIterations=int(g("nonlinearIterations")[0]),
        kappa=float(g("kappa")[0]), theta=float(g("theta")[0]), zeta=float(g("zeta")[0]),
        matrixRelaxation=float(g("matrixRelaxation")[0]), dualTimeCFL=float(g("dualTimeCFL")[0]),
        dtNondim=float(g("dtNondim")[0]), viscousCFLCoeff=float(g("viscousCFLCoeff")[0]),
        gasConstant=list(g("gasConstant")), n=list(g("n")), hf=list(g("hf")),
        nondimScaling=float(g("nondimScaling")[0]),
        suthViscC1=list(g("suthViscC1")), suthViscS=list(g("suthViscS")),
        suthCondC1=list(g("suthCondC1")), suthCondS=list(g("suthCondS")),
        molarMass=list(g("molarMass")), tRef=float(g("tRef")[0]), schmidt=float(g("schmidt")[0]),
        # sutherland::sutherland (reference src/transport.cpp:65-68): kNonDim = aRef^2 muRef / tRef
        muMixRef=float(d["cfg/muMixRef"][0]) if "cfg/muMixRef" in d else 0.0,
        kMixRef=(float(g("aRef")[0]) ** 2 * float(d["cfg/muMixRef"][0]) / float(g("tRef")[0])
                 if "cfg/muMixRef" in d else 0.0),
        bcStates=bcs)


def conns_from_dump(d, prefix=""):
    out = []
    for row in d.get(prefix + "connections", np.zeros((0, 28), dtype=np.int32)):
        c = abi.Conn()
        r = [int(v) for v in row]
        for n, name in enumerate(("rank", "block", "localBlock", "boundary", "d1Start", "d1End",
                                  "d2Start", "d2End", "constSurf")):
            arr = getattr(c, name)
            arr[0], arr[1] = r[2 * n], r[2 * n + 1]
        for q in range(8):
            c.patchBorder[q] = r[18 + q]
        c.orientation, c.isInterblock = r[26], r[27]
        out.append(c)
    return out


def multigrid_from_dump(d):
    """(problems, transfers, cycle index) of a multi-level dump: the finest level and every
    coarse level `L<l>/` as a Problem of its own, plus the transfer maps between level pairs
    (oracle.OracleMultigrid)."""
    nl = int(d["cfg/multigridLevels"][0])
    problems = [problem_from_dump(d)] + [problem_from_dump(d, prefix="L%d/" % l)
                                         for l in range(1, nl)]
    transfers = []
    for l in range(nl - 1):
        per_block = []
        for bb in range(len(problems[l].blocks)):
            p = "L%d/b%d/" % (l, bb)
            per_block.append((d[p + "toCoarse"], d[p + "volWeightFactor"], d[p + "prolongCoeffs"]))
        transfers.append(per_block)
    return problems, transfers, int(d["cfg/mgCycleIndex"][0])


def problem_from_dump(d, state_key="state0", prefix=""):
    cfg = cfg_from_dump(d)
    blocks = []
    for bb in range(int(d[prefix + "numBlocks"][0])):
        p = prefix + "b%d/" % bb
        ni, nj, nk, g, parent, rank, lpos, gpos = [int(v) for v in d[p + "dims"]]
        surfaces = [tuple(int(v) for v in row[:8]) for row in d[p + "surfaces"]]
        arrays = {k: np.array(d[p + k]) for k in ("vol", "fAreaI", "fAreaJ", "fAreaK", "center",
                                                 "cellWidthI", "cellWidthJ", "cellWidthK",
                                                 "wallDist")}
        arrays["state"] = np.array(d[p + state_key])
        blocks.append(Block(ni, nj, nk, surfaces, arrays, parent_block=parent, global_pos=gpos))
    return Problem(cfg, blocks, conns_from_dump(d, prefix))


def stage_case(src_dir, dst_dir, edits=None, iterations=None):
    """Copy a case directory (a .inp + its grid) and apply `key: value` edits to the .inp."""
    os.makedirs(dst_dir, exist_ok=True)
    inp = None
    for f in os.listdir(src_dir):
        shutil.copy(os.path.join(src_dir, f), os.path.join(dst_dir, f))
        os.chmod(os.path.join(dst_dir, f), 0o644)
        if f.endswith(".inp"):
            inp = f
    edits = dict(edits or {})
    if iterations is not None:
        edits["iterations"] = str(iterations)
    if edits:
        path = os.path.join(dst_dir, inp)
        lines = open(path).read().split("\n")
        seen = set()
        for n, line in enumerate(lines):
            key = line.split(":")[0].strip()
            if key in edits and not line.lstrip().startswith("#"):
                lines[n] = "%s: %s" % (key, edits[key])
                seen.add(key)
        extra = ["%s: %s" % (k, v) for k, v in edits.items() if k not in seen]
        # new keys must precede the boundaryConditions table
        for n, line in enumerate(lines):
            if line.startswith("boundaryConditions:") or line.startswith("boundaryStates"):
                lines[n:n] = extra
                break
        open(path, "w").write("\n".join(lines))
    return inp
