# usage: bash scripts/quick_bench.sh  -> GPU tests for EAM/MD + one bench line summary
set -e
python -m pytest tests/test_gpu_eam.py tests/test_gpu_md.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps ${STEPS:-1000} --warmup 50 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(chr(10))[-1])
n=d['steps']
print('value %.1f M atom-steps/s  ms/step %.4f  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))
print({k: round(v/n*1e3,1) for k,v in d['kernels_ms'].items() if isinstance(v, float)}, 'us per step;', d['md'])
print('roofline', d['roofline'])"
