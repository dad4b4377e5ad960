# launch list of the bench command + ncu --set full of the hot kernels (current build); the
# .ncu-rep files stay on the box (too large to bring back), their raw pages come back as csv
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --n 256 > gpurun_out/r01h_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'ImplicitTmaKernel|ResidualMarchKernel|UpdateKernel' -s 7 -c 4 -f -o /tmp/r01h_full python bench.py --steps 1 --warmup 1 --no-cpu --n 256 > gpurun_out/r01h_full.log 2>&1
ncu -i /tmp/r01h_full.ncu-rep --page raw --csv > gpurun_out/r01h_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'RansCellKernel|DplurKernel' -s 2 -c 2 -f -o /tmp/r01h_rans python bench.py --steps 1 --warmup 1 --no-cpu --n 128 --turb sst2003 > gpurun_out/r01h_rans.log 2>&1
ncu -i /tmp/r01h_rans.ncu-rep --page raw --csv > gpurun_out/r01h_rans_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r01h_full_raw.csv > gpurun_out/r01h_ncu_full_summary.txt 2>&1
python scripts/ncu_summary.py gpurun_out/r01h_rans_raw.csv > gpurun_out/r01h_ncu_rans_summary.txt 2>&1
ls -la gpurun_out | tail -12
