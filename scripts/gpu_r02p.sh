# round 2, step p: (2 GPUs) whole -m gpu suite incl. the NCCL parity tests; bench with the overlapped exchange vs serial
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02p_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError:|passed|failed" gpurun_out/r02p_pytest_gpu.txt | head -30
run2() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" > gpurun_out/r02p_$name.json 2> gpurun_out/r02p_$name.err || tail -5 gpurun_out/r02p_$name.err; }
AITHER_BENCH_CONFIGS3_DIMS=256,128,128 run2 overlap2
AITHER_B200_HALO_OVERLAP=0 run2 serial2 --no-configs3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r02p_one.json 2> gpurun_out/r02p_one.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('r02p_')[1][:-5], 'n', d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'exch', d.get('exchange_ms_per_step'), d['kernel_ms_per_step'])
        if 'configs3' in d: print('   configs3', {k:v for k,v in d['configs3'].items() if k!='workload'})
    except Exception as e:
        print(f, 'failed', e)
PY
