cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 25 python -m pytest tests/test_gpu_multigrid.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r01r_pytest_gpu_multigrid.txt
