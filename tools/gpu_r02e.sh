#!/bin/bash
# round-2 visit E: k_sdf_fwd16 in front of the alpha cull, vectorised smem loads in k_corr_persist
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/${TAG}_parity.log
for f in 1 0; do
ARAH_SDF_FWD16=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline > gpurun_out/${TAG}_bench_f$f.json 2> gpurun_out/${TAG}_bench_f$f.err; echo "bench fwd16=$f rc=$?"
done
TAG=$TAG python - <<'PY'
import json, os
for f in (1, 0):
    try:
        d=json.loads(open('gpurun_out/%s_bench_f%d.json' % (os.environ['TAG'], f)).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d.get('stages_ms_last_step').items()})
        print(d['phase_cycles_last_step']['corr'])
    except Exception as e:
        print(f, 'failed', e)
PY
