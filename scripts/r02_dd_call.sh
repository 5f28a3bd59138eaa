#!/bin/bash
# 2-GPU call: decomposition parity tests (peer-to-peer and NCCL step paths) + bench at N=2 with the
# rebuild phase timers.  gpurun --gpus 2 --timeout 1200 -- 'bash scripts/r02_dd_call.sh'
set -u
OUT=gpurun_out/r02_dd
mkdir -p $OUT
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi -L | tee $OUT/summary.txt
echo "== neighbour list parity with the single-precision pre-filter" | tee -a $OUT/summary.txt
timeout 600 python -m pytest tests/test_gpu_neighbors.py -m gpu -q > $OUT/pytest_nl.log 2>&1
echo "exit $?" | tee -a $OUT/summary.txt; tail -15 $OUT/pytest_nl.log | tee -a $OUT/summary.txt
export ATX_NL_F32=0     # the decomposition runs of this call use the verified pair search
echo "== DD parity tests, p2p step path" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests/test_gpu_dd.py tests/test_gpu_multidevice.py -m gpu -q -x > $OUT/pytest_p2p.log 2>&1
echo "exit $?" | tee -a $OUT/summary.txt; tail -15 $OUT/pytest_p2p.log | tee -a $OUT/summary.txt
echo "== DD parity tests, NCCL step path" | tee -a $OUT/summary.txt
ATX_DD_P2P=0 timeout 600 python -m pytest tests/test_gpu_dd.py -m gpu -q -x -k "eam or tersoff" > $OUT/pytest_nccl.log 2>&1
echo "exit $?" | tee -a $OUT/summary.txt; tail -5 $OUT/pytest_nccl.log | tee -a $OUT/summary.txt
echo "== bench N=2 (driver flags), p2p" | tee -a $OUT/summary.txt
timeout 900 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 --blocks c2,c4 > $OUT/bench_n2_p2p.json 2> $OUT/bench_n2_p2p.err
echo "exit $?" | tee -a $OUT/summary.txt; tail -c 6000 $OUT/bench_n2_p2p.json | tee -a $OUT/summary.txt; tail -5 $OUT/bench_n2_p2p.err
echo "== bench N=2, p2p, exact rebuild phases" | tee -a $OUT/summary.txt
ATX_DD_PROFILE=1 timeout 900 $TR --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --blocks c2,c4 --c4-steps 40 > $OUT/bench_n2_prof.json 2> $OUT/bench_n2_prof.err
echo "exit $?" | tee -a $OUT/summary.txt
echo "== bench N=2, NCCL step path" | tee -a $OUT/summary.txt
ATX_DD_P2P=0 timeout 900 $TR --master-port 29543 bench.py --gpus 2 --steps 20 --warmup 5 --blocks c2,c4 > $OUT/bench_n2_nccl.json 2> $OUT/bench_n2_nccl.err
echo "exit $?" | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
for n in ('p2p', 'prof', 'nccl'):
    try:
        d = json.loads(open('gpurun_out/r02_dd/bench_n2_%s.json' % n).read().strip().split('\n')[-1])
        print(n, 'C2 value %.1f M, ms/step %.4f, rebuilds %d, steady %.1f M; dd %s' % (
            d['value'] / 1e6, d['ms_per_step'], d['md']['rebuilds'], d['steady_state']['value'] / 1e6, d['kernels_ms']['dd']))
        print('   parity', d.get('parity'))
        for k, b in (d.get('c4') or {}).items():
            print('   C4', k, '%.1f M, ms/step %.4f, halo/step %.4f, bop/step %.4f' % (b['value'] / 1e6, b['ms_per_step'], b['dd_halo_ms_per_step_rank0'], b['bop_force_ms_per_step_rank0']), b.get('rebuild_host_ms_since_create'), b.get('parity'))
        print('   errors', d.get('block_errors'))
    except Exception as e:
        print(n, 'failed', e)
PY
