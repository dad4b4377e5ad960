cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multiblock.py -m gpu -x -q -k "couette or rae2822" 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_two.txt
