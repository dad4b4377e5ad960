set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r1_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --n 256 > gpurun_out/r1_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'MarchKernel' -s 6 -c 4 -o gpurun_out/r1_march python bench.py --steps 1 --warmup 1 --no-cpu --n 256 > gpurun_out/r1_ncu.log 2>&1
ls -la gpurun_out
