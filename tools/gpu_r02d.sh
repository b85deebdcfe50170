#!/bin/bash
# round-2 visit D: per-point phase of k_corr_persist optimised, phase clocks of k_trace_persist
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/${TAG}_parity.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
TAG=$TAG python - <<'PY'
import json, os
d=json.loads(open('gpurun_out/%s_bench.json' % os.environ['TAG']).read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d.get('stages_ms_last_step').items()})
print(d['phase_cycles_last_step'])
PY
