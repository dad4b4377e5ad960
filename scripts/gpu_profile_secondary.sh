# ncu --set full of the secondary-path kernels (RANS cell pass, block diagonal / inverse, split LU-SGS
# plane kernel, cell-parallel block DPLUR); raw pages come back as csv summaries
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:'RansCellKernel|BlockDiagInvKernel|PrepBlockKernel' -s 3 -c 3 -f -o /tmp/sec_a python bench.py --steps 1 --warmup 1 --no-cpu --n 96 --turb sst2003 --solver bdplur > gpurun_out/sec_a.log 2>&1
ncu -i /tmp/sec_a.ncu-rep --page raw --csv > /tmp/sec_a.csv 2>/dev/null; python scripts/ncu_summary.py /tmp/sec_a.csv > gpurun_out/r01m_ncu_rans_block_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'LusgsPlaneSplitKernel' -s 400 -c 2 -f -o /tmp/sec_b python bench.py --steps 1 --warmup 1 --no-cpu --n 128 --solver lusgs > gpurun_out/sec_b.log 2>&1
ncu -i /tmp/sec_b.ncu-rep --page raw --csv > /tmp/sec_b.csv 2>/dev/null; python scripts/ncu_summary.py /tmp/sec_b.csv > gpurun_out/r01m_ncu_lusgs_summary.txt 2>&1
wc -l gpurun_out/r01m_ncu_*_summary.txt; tail -n 3 gpurun_out/sec_a.log; tail -n 3 gpurun_out/sec_b.log
