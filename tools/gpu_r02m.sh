#!/bin/bash
# shade16 bring-up: parity with the fp16 shading kernel, then A/B bench against k_shade_tc3
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for k in 1 0; do
ARAH_SHADE16=$k timeout 600 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench_s$k.json 2> gpurun_out/${TAG}_bench_s$k.err; echo "bench shade16=$k rc=$?"
K=$k TAG=$TAG python - <<'PY'
import json, os
k, t = os.environ['K'], os.environ['TAG']
try:
    d=json.loads(open(f'gpurun_out/{t}_bench_s{k}.json').read().strip().splitlines()[-1])
    print(k, round(d['value']), round(d['ms_per_step'],2), {a: round(v,2) for a,v in d['stages_ms_last_step'].items()}, d.get('parity'))
    print(d['phase_cycles_last_step'].get('shade'))
except Exception as e:
    print('parse failed', e)
PY
done
