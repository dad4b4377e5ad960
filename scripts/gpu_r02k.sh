# round 2, step k: plane-major workspace, bulk-copy ring; scaling probe of the half sweep
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python scripts/probe_lusgs.py gpurun_out/r02k_probe.json 2>&1 | tail -12
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or phases or rans or multiblock" > gpurun_out/r02k_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02k_pytest_gpu.txt
