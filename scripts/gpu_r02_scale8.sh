# round 2: N-GPU bench line (weak scaling, with configs3) -- N from the number of visible GPUs
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
S=$SECONDS; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 $BENCH_EXTRA > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err; echo "bench$N rc=$? $((SECONDS-S))s"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_scale_${N}gpu.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','exchange_ms_per_step')}); print('e2e',d['e2e']['value'], d['kernel_ms_per_step'])
c=d.get('configs3'); print({k:c.get(k) for k in ('ms_per_step','ms_per_step_1gpu_same_run','value','exchange_ms_per_step')} if c else None)
PY
