#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['stages_ms_last_step'].items()})
"
bash tools/gpu_ncu_all.sh ${TAG}
