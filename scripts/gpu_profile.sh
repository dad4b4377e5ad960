# ncu evidence for profiles/: launch list of one bench step + full capture of the hot kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/$1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/$1_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ResidualMarchKernel|ImplicitTmaKernel|UpdateKernel' -s 3 -c 4 -f -o gpurun_out/$1_full python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/$1_full.log 2>&1
tail -2 gpurun_out/$1_full.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err; tail -2 gpurun_out/$1_bench.err
bash scripts/gpu_variants.sh
