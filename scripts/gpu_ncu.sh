# usage: gpu_ncu.sh <kernel-regex> <skip> <count> <outname>
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 python bench.py --steps 1 --warmup 1 --no-cpu --n 256 > gpurun_out/$4.log 2>&1
tail -3 gpurun_out/$4.log
