#!/bin/bash
# round-2 visit C: persistent tracing / joint-search kernels: parity in isolation, then together, then A/B timing
TAG=${1:-r02c}
mkdir -p gpurun_out
ARAH_TRACE_PERSIST=1 ARAH_ISO_PERSIST=0 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "golden and tf32" > gpurun_out/${TAG}_parity_trace.log 2>&1; echo "parity trace-only rc=$?"
tail -3 gpurun_out/${TAG}_parity_trace.log
ARAH_TRACE_PERSIST=0 ARAH_ISO_PERSIST=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "golden and tf32" > gpurun_out/${TAG}_parity_iso.log 2>&1; echo "parity iso-only rc=$?"
tail -3 gpurun_out/${TAG}_parity_iso.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_parity.log 2>&1; echo "parity all rc=$?"
tail -3 gpurun_out/${TAG}_parity.log
for cfg in "1 1" "0 0" "1 0" "0 1"; do
  set -- $cfg
  ARAH_TRACE_PERSIST=$1 ARAH_ISO_PERSIST=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline > gpurun_out/${TAG}_bench_t$1_i$2.json 2> gpurun_out/${TAG}_bench_t$1_i$2.err; echo "bench t$1 i$2 rc=$?"
done
TAG=$TAG python - <<'PY'
import json, os
for t in ('t1_i1','t0_i0','t1_i0','t0_i1'):
    try:
        d=json.loads(open('gpurun_out/%s_bench_%s.json' % (os.environ['TAG'], t)).read().strip().splitlines()[-1])
        print(t, round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d.get('stages_ms_last_step').items()}, d.get('counters_last_step'))
    except Exception as e:
        print(t, 'failed', e)
PY
