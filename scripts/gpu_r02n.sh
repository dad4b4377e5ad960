# round 2, step n: whole -m gpu suite with the stricter checker (own-scale per component), the
# multi-tile oracle tests, multigrid without xfail; subsonicCylinder / multiblockCylinder bisect
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError:|passed|failed" gpurun_out/r02n_pytest_gpu.txt | head -60
timeout 300 python scripts/diag_subsonic.py subsonicCylinder 2>&1 | tee gpurun_out/r02n_diag_subsonic.txt
