# BASELINE configs[3]'s scheme (WENO5 + viscous centralFourth, DPLUR, halo exchange) at 1..N GPUs
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
true
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --n 192 --recon weno --viscous > gpurun_out/c4_1.json 2> gpurun_out/c4_1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $n --steps 5 --warmup 3 --cells 192 --recon weno --viscous > gpurun_out/c4_$n.json 2> gpurun_out/c4_$n.err || tail -5 gpurun_out/c4_$n.err
  fi
done
python - <<'PY'
import json
base=None
for n in (1,2,4,8):
    try:
        d=json.loads([l for l in open('gpurun_out/c4_%d.json'%n) if l.startswith('{')][-1])
    except Exception as e:
        continue
    if n==1: base=d['value']
    print(n, 'ms/step %.3f'%d['ms_per_step'], 'value %.0f'%d['value'], 'eff %.3f'%(d['value']/n/base if base else 0), d['kernel_ms_per_step'])
PY
