#!/bin/bash
# banded lattice + shade16 + softplus mix: mesh and parity tests, then a bench with the mesh entry
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_mesh.py tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -a "banded lattice\|passed\|failed\|Error\|error" gpurun_out/${TAG}_pytest.log | head -30
timeout 900 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
TAG=$TAG python - <<'PY'
import json, os
t = os.environ['TAG']
try:
    d=json.loads(open(f'gpurun_out/{t}_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],2), {a: round(v,2) for a,v in d['stages_ms_last_step'].items()})
    print(d['phase_cycles_last_step'].get('corr'))
    print('mesh', json.dumps(d.get('mesh_extract'))[:1500])
except Exception as e:
    print('parse failed', e)
PY
tail -5 gpurun_out/${TAG}_bench.err
