# launch list of the bench command + ncu --set full of the two marching kernels (current build)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --n 256 > gpurun_out/r01h_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ImplicitTmaKernel|ResidualMarchKernel|UpdateKernel' -s 7 -c 4 -f -o gpurun_out/r01h_full python bench.py --steps 1 --warmup 1 --no-cpu --n 256 > gpurun_out/r01h_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'RansCellKernel|DplurKernel' -s 2 -c 2 -f -o gpurun_out/r01h_rans python bench.py --steps 1 --warmup 1 --no-cpu --n 128 --turb sst2003 > gpurun_out/r01h_rans.log 2>&1
ls -la gpurun_out | tail -8
