# round 2, step u: (2 GPUs) the bench line as the driver runs it at N = 2, configs[3] block at full size (512 x 256 x 256 per GPU)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; SECONDS=0
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02u_scale2.json 2> gpurun_out/r02u_scale2.err || tail -5 gpurun_out/r02u_scale2.err

echo "bench wall seconds: $SECONDS"; free -g | head -2; nproc
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02u_scale2.json') if l.startswith('{')][-1])
print('n', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'value %.0f'%d['value'], 'exch', d.get('exchange_ms_per_step'), d['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']))
print('configs3', {k:v for k,v in d['configs3'].items() if k!='workload'})
PY
