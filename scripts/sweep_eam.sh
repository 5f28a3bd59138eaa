# sweep of the EAM fast-kernel launch shape (lanes per atom x list entries in flight per lane)
set -e
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for cfg in ${CFGS:-"16 1" "8 2" "8 4" "4 2" "4 4" "2 4"}; do
  set -- $cfg
  echo "LANES=$1 UNROLL=$2"
  ATX_EAM_LANES=$1 ATX_EAM_UNROLL=$2 python bench.py --steps 300 --warmup 20 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('  value %.1f M atom-steps/s  ms/step %.3f  force %.1f us  density %.1f us  e2e %.1f M' % (d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms']*1e3, d['kernels_ms']['eam_density']/max(d['roofline']['launches'],1)*1e3, d['e2e']['value']/1e6))"
done
