#!/bin/bash
# Quick A/B of one env switch on the GPU box: parity tests with the new setting, then the bench line with the switch on and off.
# Usage: tools/gpu_ab.sh <tag> <ENV_NAME>
TAG=${1:-ab}; VAR=${2:-ARAH_CORR_INTERLEAVE}
mkdir -p gpurun_out
timeout -s KILL 420 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for v in 1 0; do
  env ${VAR}=$v timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench ${VAR}=$v rc=$?"
  python - <<EOF
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_$v.json').read().splitlines()[0])
    print('${VAR}=$v', round(d['value']), 'rays/s', round(d['ms_per_step'],2), 'ms', d['stages_ms_last_step'])
    print(' corr phases', d['phase_cycles_last_step']['corr'])
except Exception as e:
    print('no bench line', e)
EOF
done
