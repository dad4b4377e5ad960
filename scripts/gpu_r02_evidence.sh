# round 2 evidence of the final build: launch list of the bench command, ncu --set full of the hot
# kernels (-> summaries + profiles/traffic.json source), LU-SGS pencil kernel, compute-sanitizer
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs3 > gpurun_out/r02_launches.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'ImplicitTmaKernel|ResidualMarchKernel|UpdateKernel' -s 7 -c 7 -f -o /tmp/r02_full python bench.py --steps 1 --warmup 1 --no-cpu --no-configs3 > gpurun_out/r02_full.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_full_raw.csv > gpurun_out/r02_ncu_full_summary.txt 2>&1
python scripts/make_traffic.py gpurun_out/r02_full_raw.csv gpurun_out/r02_traffic.json "profiles/r02_ncu_full_summary.txt (ncu --set full of the final round-2 build, scripts/gpu_r02_evidence.sh)" > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:LusgsPencilKernel -s 4 -c 2 -f -o /tmp/r02_pencil python bench.py --steps 1 --warmup 1 --no-cpu --no-configs3 --n 192 --solver lusgs > gpurun_out/r02_pencil.log 2>&1
ncu -i /tmp/r02_pencil.ncu-rep --page raw --csv > gpurun_out/r02_pencil_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_pencil_raw.csv > gpurun_out/r02_ncu_lusgs_pencil_summary.txt 2>&1
python scripts/make_traffic.py gpurun_out/r02_pencil_raw.csv gpurun_out/r02_traffic_lusgs.json "profiles/r02_ncu_lusgs_pencil_summary.txt (192^3)" 7077888 > /dev/null 2>&1
rm -f gpurun_out/r02_full_raw.csv gpurun_out/r02_pencil_raw.csv
bash scripts/gpu_sanitize.sh > gpurun_out/r02_sanitize_tail.txt 2>&1; cp gpurun_out/sanitize.log gpurun_out/r02_compute_sanitizer_memcheck.log
tail -5 gpurun_out/r02_sanitize_tail.txt; cat gpurun_out/r02_traffic.json | head -30
