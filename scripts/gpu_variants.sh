# secondary workloads (not the headline): WENO5 inviscid, and WENO5 + viscous at 192^3
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for v in "--recon weno" "--recon weno --viscous" "--viscous"; do
  timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --n 192 $v > gpurun_out/variant.json 2> gpurun_out/variant.err || tail -3 gpurun_out/variant.err
  python - <<PY
import json
d=json.load(open('gpurun_out/variant.json'))
print("$v", 'ms/step %.2f' % d['ms_per_step'], 'Mcell-iter/s %.0f' % d['value'], d['kernel_ms_per_step'])
PY
done
