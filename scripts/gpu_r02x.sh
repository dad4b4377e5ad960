# round 2, step x: ncu --set full of the one-thread-per-cell LU-SGS pencil kernel at 192^3
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:LusgsPencilKernel -s 4 -c 2 -f -o gpurun_out/r02x_lusgs_pencil python bench.py --steps 1 --warmup 1 --no-cpu --no-configs3 --n 192 --solver lusgs > gpurun_out/r02x_ncu.log 2>&1
tail -3 gpurun_out/r02x_ncu.log
