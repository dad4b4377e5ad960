# round 2, step o: (1 GPU) whole -m gpu suite after the checker / overlap changes; chunk-length sweep of the headline
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02o_pytest_gpu.txt 2>&1; grep -E "^FAILED|^ERROR|AssertionError:|passed|failed" gpurun_out/r02o_pytest_gpu.txt | head -40
timeout 300 python scripts/diag_subsonic.py subsonicCylinder 2>&1 | tee gpurun_out/r02o_diag_subsonic.txt
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/r02o_$name.json 2> gpurun_out/r02o_$name.err || tail -3 gpurun_out/r02o_$name.err; }
run base
for c in 16 24 40 48 64; do AITHER_B200_RES_CHUNK=$c run res$c; done
for c in 16 22 26 43 64; do AITHER_B200_TMA_CHUNK=$c run tma$c; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o_*.json')):
    try:
        d=json.load(open(f))
        print(f.split('r02o_')[1][:-5], 'ms/step %.3f' % d['ms_per_step'], d['kernel_ms_per_step'])
    except Exception as e:
        print(f, 'failed', e)
PY
