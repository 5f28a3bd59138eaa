#!/bin/bash
# 8-GPU call: decomposition parity at 4 ranks (interior ranks with two distinct neighbours) and the
# driver-flag bench at N=8.  gpurun --gpus 8 --timeout 900 -- 'bash scripts/r02_n8_call.sh'
set -u
OUT=gpurun_out/r02_n8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l | tee $OUT/summary.txt
timeout 300 python -m pytest tests/test_gpu_dd.py -m gpu -q -x -k "eam or tersoff" > $OUT/pytest_dd4.log 2>&1
echo "dd tests (4 ranks) exit $?" | tee -a $OUT/summary.txt; tail -3 $OUT/pytest_dd4.log | tee -a $OUT/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/bench_n8.json 2> $OUT/bench_n8.err
echo "bench n8 exit $?" | tee -a $OUT/summary.txt; tail -3 $OUT/bench_n8.err | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
try:
    d = json.loads(open('gpurun_out/r02_n8/bench_n8.json').read().strip().split('\n')[-1])
    print('C2 N=8 value %.1f M, ms/step %.4f, rebuilds %d, steady %.1f M (%.4f ms/step); e2e %.1f M' % (
        d['value'] / 1e6, d['ms_per_step'], d['md']['rebuilds'], d['steady_state']['value'] / 1e6, d['steady_state']['ms_per_step'], d['e2e']['value'] / 1e6))
    print('   dd', d['kernels_ms']['dd'], 'force', d['kernels_ms']['eam_force'], 'dens', d['kernels_ms']['eam_density'])
    print('   parity', d.get('parity'))
    for k, b in (d.get('c4') or {}).items():
        print('   C4', k, '%.1f M, ms/step %.4f, halo/step %.4f, bop/step %.4f rebuilds %d ghosts %d' % (b['value'] / 1e6, b['ms_per_step'], b['dd_halo_ms_per_step_rank0'], b['bop_force_ms_per_step_rank0'], b['rebuilds'], b['ghost_atoms_rank0']), b.get('parity'))
    print('   errors', d.get('block_errors'), d.get('clocks'))
except Exception as e:
    print('failed', e)
PY
