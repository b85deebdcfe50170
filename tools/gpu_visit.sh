#!/bin/bash
# One GPU-box visit as the driver does it at round end: all GPU tests, the default bench line, the reference arm.
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -a "passed\|failed\|Error\|error" gpurun_out/${TAG}_pytest.log | tail -5
timeout 1500 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc=$?"
TAG=$TAG python - <<'PY'
import json, os
t = os.environ['TAG']
try:
    d=json.loads(open(f'gpurun_out/{t}_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {a: round(v,2) for a,v in d['stages_ms_last_step'].items()}, d['gpu_launches_per_frame'])
    for k in ('sequence258', 'h36m_1024', 'strict_fp32', 'train_step', 'mesh_extract', 'hypernet', 'ray_setup', 'image_tail', 'parity', 'cpu_baseline', 'roofline'):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print('parse failed', e)
try:
    d=json.loads(open(f'gpurun_out/{t}_bench_ref.json').read().strip().splitlines()[-1])
    print('ref', d['value'], d['cpu_baseline'])
except Exception as e:
    print('ref parse failed', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
