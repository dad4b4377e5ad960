# round 2 final, two GPUs: the 2-GPU tests, smoke with two ranks, the N = 2 bench line (with configs3)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_multiblock.py -m gpu -q > gpurun_out/r02_pytest_2gpu.txt 2>&1; tail -3 gpurun_out/r02_pytest_2gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_2gpu.txt 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r02_smoke_2gpu.txt
S=$SECONDS; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_scale_2gpu.json 2> gpurun_out/r02_scale_2gpu.err; echo "bench2 rc=$? $((SECONDS-S))s"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_scale_2gpu.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','exchange_ms_per_step','configs3')}); print(d['e2e']['value'], d['kernel_ms_per_step'])
PY
