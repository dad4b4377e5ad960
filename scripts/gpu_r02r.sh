# round 2, step r: (2 GPUs) whole -m gpu suite incl. the shim on the shipped cases and the NCCL parity tests; one-layer update exchange
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02r_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError|passed|failed" gpurun_out/r02r_pytest_gpu.txt | head -30
run2() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/r02r_$name.json 2> gpurun_out/r02r_$name.err || tail -5 gpurun_out/r02r_$name.err; }
run2 onelayer2 --no-configs3
AITHER_B200_HALO_LAYERS=all run2 alllayers2 --no-configs3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02r_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('r02r_')[1][:-5], 'n', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'exch', d.get('exchange_ms_per_step'), d['kernel_ms_per_step'], 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(f, 'failed', e)
PY
