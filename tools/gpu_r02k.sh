#!/bin/bash
TAG=${1:-r02k}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -s -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
for k in 2 1; do
ARAH_KNN_SEED=$k timeout 600 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench_k$k.json 2> gpurun_out/${TAG}_bench_k$k.err; echo "bench knn_seed=$k rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_k$k.json').read().strip().splitlines()[-1])
print($k, round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['stages_ms_last_step'].items()})
"
done
