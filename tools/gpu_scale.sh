#!/bin/bash
# multi-GPU bench line as the driver launches it: tools/gpu_scale.sh <N> <tag>
N=${1:-2}
TAG=${2:-r02s}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/${TAG}_scale$N.json 2> gpurun_out/${TAG}_scale$N.err; echo "bench N=$N rc=$?"
tail -c 600 gpurun_out/${TAG}_scale$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
    > gpurun_out/${TAG}_scale${N}_ref.json 2> gpurun_out/${TAG}_scale${N}_ref.err; echo "ref N=$N rc=$?"
N=$N TAG=$TAG python - <<'PY'
import json, os
n, t = os.environ['N'], os.environ['TAG']
try:
    d=json.loads(open(f'gpurun_out/{t}_scale{n}.json').read().strip().splitlines()[-1])
    print(n, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['per_rank_ms_per_step'])
    print('sequence258', json.dumps(d.get('sequence258'))[:900])
except Exception as e:
    print('parse failed', e)
try:
    d=json.loads(open(f'gpurun_out/{t}_scale{n}_ref.json').read().strip().splitlines()[-1])
    print('ref', d['value'], d['cpu_baseline']['cores'])
except Exception as e:
    print('ref parse failed', e)
PY
