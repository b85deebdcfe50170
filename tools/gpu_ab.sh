#!/bin/bash
# A/B of one library switch: tools/gpu_ab.sh <tag> <ENV_NAME> [values...]   (default values: 1 0)
TAG=${1:-ab}
NAME=${2:-ARAH_RAY_ORDER}
shift 2
VALS=${@:-1 0}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -a "passed\|failed\|Error\|error" gpurun_out/${TAG}_pytest.log | tail -5
for v in $VALS; do
env $NAME=$v timeout 900 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench_$v.json 2> gpurun_out/${TAG}_bench_$v.err; echo "bench $NAME=$v rc=$?"
V=$v TAG=$TAG python - <<'PY'
import json, os
v, t = os.environ['V'], os.environ['TAG']
try:
    d=json.loads(open(f'gpurun_out/{t}_bench_{v}.json').read().strip().splitlines()[-1])
    print(v, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a: round(x,2) for a,x in d['stages_ms_last_step'].items()}, d['gpu_launches_per_frame'])
    print('  trace phases', d['phase_cycles_last_step'].get('trace'))
except Exception as e:
    print('parse failed', e)
PY
tail -2 gpurun_out/${TAG}_bench_$v.err
done
