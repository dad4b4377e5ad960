# round 2, step c: ncu source-level profile of the pencil LU-SGS kernel (128^3, one fullGS launch)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:LusgsPencil -s 3 -c 1 -o gpurun_out/r02c_pencil -f python bench.py --n 128 --solver lusgs --steps 1 --warmup 1 --no-cpu > gpurun_out/r02c_ncu.log 2>&1; tail -3 gpurun_out/r02c_ncu.log
