# round 2, step w: one-thread-per-cell LU-SGS pencil kernel: parity first, then timing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_phases.py tests/test_gpu_multiblock.py tests/test_gpu_rans.py tests/test_gpu_viscous.py -m gpu -q -x -k "lusgs or subsonicCylinder or viscousFlatPlate or turbFlatPlate or box_kw or uniformFlow or inlet_outlet or periodic or multiblock or shockTube or LUSGS or Lusgs" > gpurun_out/r02w_pytest_lusgs.txt 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02w_pytest_lusgs.txt
timeout 300 python scripts/probe_lusgs.py > gpurun_out/r02w_lusgs_probe.json 2> gpurun_out/r02w_probe.err; echo "probe rc=$?"; cat gpurun_out/r02w_lusgs_probe.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --n 192 --solver lusgs > gpurun_out/r02w_lusgs192.json 2> gpurun_out/r02w_lusgs192.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02w_lusgs192.json
