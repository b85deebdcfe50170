#!/bin/bash
# round-2 visit H: quad-cooperative 1-NN (trace kernel + k_knn_samples), then ncu captures of every kernel
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for sd in 0 1; do
ARAH_KNN_SEED=$sd timeout 600 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench_s$sd.json 2> gpurun_out/${TAG}_bench_s$sd.err; echo "bench seed=$sd rc=$?"
done
TAG=$TAG python - <<'PY'
import json, os
t=os.environ['TAG']
for sd in (0,1):
    try:
        d=json.loads(open('gpurun_out/%s_bench_s%d.json' % (t, sd)).read().strip().splitlines()[-1])
        print(sd, round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['stages_ms_last_step'].items()})
        print(d['phase_cycles_last_step']['trace'])
    except Exception as e:
        print('bench parse failed', e)
PY
bash tools/gpu_ncu_all.sh ${TAG}
