cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for rc in 0 32 43 64 86 128; do for tc in 0 32 43 64 128; do
  if [ $rc != 0 ] && [ $tc != 0 ] && [ $rc != $tc ]; then continue; fi
  if [ $rc != 0 ]; then export AITHER_B200_RES_CHUNK=$rc; else unset AITHER_B200_RES_CHUNK; fi
  if [ $tc != 0 ]; then export AITHER_B200_TMA_CHUNK=$tc; else unset AITHER_B200_TMA_CHUNK; fi
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/chunk.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/chunk.json')); k=d['kernel_ms_per_step']
print('res_chunk $rc tma_chunk $tc  ms/step %.3f residual %.3f dplur %.3f axmb %.3f' % (d['ms_per_step'], k['residual'], k['dplur_sweep'], k['matrix_residual']))
PY
done; done
