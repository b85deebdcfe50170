#!/bin/bash
# round-2 visit G: 16-warp SDF engine in k_trace_persist / k_iso_persist / k_sdf_grid16
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py tests/test_gpu_train.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --seq-frames 8 --no-h36m > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/${TAG}_bench.err
TAG=$TAG python - <<'PY'
import json, os
t=os.environ['TAG']
try:
    d=json.loads(open('gpurun_out/%s_bench.json' % t).read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['stages_ms_last_step'].items()})
    print(d['phase_cycles_last_step']['trace'])
    for k in ('sequence258','mesh_extract'):
        print(k, json.dumps(d.get(k))[:900])
except Exception as e:
    print('bench parse failed', e)
PY
