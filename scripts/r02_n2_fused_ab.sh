#!/bin/bash
# 2 GPUs: decomposition tests, then C2 + C4 (80^3 cells: the per-rank size of the 8-GPU run) with the
# leapfrog-form step on and off
set -u
OUT=gpurun_out/r02_n2ab
mkdir -p $OUT
export PYTHONUNBUFFERED=1
python -m pytest tests/test_gpu_dd.py -m gpu -q 2>&1 | tail -5 | tee $OUT/pytest.log
ATX_DD_P2P=0 python -m pytest tests/test_gpu_dd.py -m gpu -q 2>&1 | tail -3 | tee -a $OUT/pytest.log
for fused in 1 0; do
ATX_DD_FUSED=$fused ATX_MD_FUSED=$fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 --blocks c4 --c4-kinds Tersoff --c4-steps 120 --c4-cells 80 > $OUT/bench_f$fused.json 2> $OUT/bench_f$fused.err
echo "exit $?" | tee -a $OUT/summary.txt
python - $fused <<'PY' | tee -a $OUT/summary.txt
import json, sys
d = json.loads(open('gpurun_out/r02_n2ab/bench_f%s.json' % sys.argv[1]).read().strip().split('\n')[-1])
print('fused', sys.argv[1], 'C2 %.1f M %.4f ms steady %.1f M parity %s' % (d['value'] / 1e6, d['ms_per_step'], d['steady_state']['value'] / 1e6, d.get('parity', {}).get('ok')), d.get('parity'))
for k, b in d['c4'].items():
    print('fused', sys.argv[1], 'C4', k, '%.1f M ms/step %.4f rebuilds %d' % (b['value'] / 1e6, b['ms_per_step'], b['rebuilds']), b.get('scopes_ms_per_step_rank0'), b.get('parity'))
PY
done
