# round 2, last check of the final commit: whole GPU suite (-x as the driver runs it), one small LU-SGS bench line
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02_last_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02_last_pytest_gpu.txt
timeout 120 python bench.py --n 96 --solver lusgs --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline']['algorithmic_bytes_per_launch'])"
