"""Scaling probe of the LU-SGS half sweep: time per launch for blocks of different shapes
(one pencil, one row of pencils, a full lattice) -- separates the per-plane latency of a pencil
from the lag between dependent pencils and from throughput."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aither_b200
from aither_b200 import synthetic

shapes = [(128, 8, 8), (512, 8, 8), (128, 64, 8), (128, 8, 64), (128, 128, 8), (128, 64, 64),
          (128, 128, 128), (192, 192, 192)]
out = []
for ni, nj, nk in shapes:
    prob = synthetic.box_problem(ni, nj, nk, solver="lusgs", sweeps=2)
    gl = aither_b200.GridLevel(prob, device=0)
    for it in range(2):
        gl.store_old_solution(it); gl.iterate(50.0)
    gl.profile_enable(True)
    n = 3
    for it in range(n):
        gl.store_old_solution(it + 2); gl.iterate(50.0)
    pr = gl.profile()
    ms, launches = pr["lusgs_plane"]
    ahead_ms = pr.get("lusgs_ahead", (0, 1))
    planes = ni + nj + nk - 2
    rec = dict(shape=[ni, nj, nk], ms_per_half_sweep=ms / launches, launches=launches,
               us_per_plane=1e3 * ms / launches / planes, cells_per_ns=ni * nj * nk / (ms / launches) * 1e-6,
               pack_ms=pr.get("lusgs_pack", (0, 1))[0] / max(1, pr.get("lusgs_pack", (0, 1))[1]),
               ahead_ms=ahead_ms[0] / max(1, ahead_ms[1]))
    print(json.dumps(rec)); out.append(rec)
    gl.close()
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/probe_lusgs.json", "w"), indent=1)
