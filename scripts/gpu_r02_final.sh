# round 2 final: default bench line (with the CPU sample), the reference arm, smoke
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
S=$SECONDS; timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$? $((SECONDS-S))s"
S=$SECONDS; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference rc=$? $((SECONDS-S))s"
S=$SECONDS; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.txt 2>&1; echo "smoke rc=$? $((SECONDS-S))s"; tail -3 gpurun_out/r02_smoke.txt
python - <<'PY'
import json
for f in ('r02_bench_final','r02_bench_reference'):
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
    print(f, {k:d.get(k) for k in ('value','unit','ms_per_step','e2e','roofline','cpu_baseline','gpu_launches','clocks')})
    if 'kernel_ms_per_step' in d: print(d['kernel_ms_per_step'], d.get('roofline_iteration'))
PY
