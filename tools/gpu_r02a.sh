#!/bin/bash
# round-2 visit A: fp16 split-precision tile probe, persistent correspondence kernel parity + A/B timing
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_00_umma.py -q -s > gpurun_out/${TAG}_probe.log 2>&1; echo "probe rc=$?" | tee -a gpurun_out/${TAG}_probe.log
tail -5 gpurun_out/${TAG}_probe.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/${TAG}_parity.log
tail -5 gpurun_out/${TAG}_parity.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline > gpurun_out/${TAG}_bench_persist.json 2> gpurun_out/${TAG}_bench_persist.err; echo "bench persist rc=$?"
ARAH_CORR_PERSIST=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline > gpurun_out/${TAG}_bench_tc5.json 2> gpurun_out/${TAG}_bench_tc5.err; echo "bench tc5 rc=$?"
TAG=$TAG python - <<'PY'
import json, os
for t in ('persist','tc5'):
    try:
        d=json.loads(open('gpurun_out/%s_bench_%s.json' % (os.environ['TAG'], t)).read().strip().splitlines()[-1])
        print(t, d['value'], d['ms_per_step'], d.get('stages_ms_last_step'), d.get('counters_last_step'))
    except Exception as e:
        print(t, 'failed', e)
PY
