#!/bin/bash
# round-2 visit F: whole GPU suite after the clean-up, full default bench line, reference arm, ncu launch list
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${TAG}_bench.err
OMP_NUM_THREADS=1 timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
TAG=$TAG python - <<'PY'
import json, os
t=os.environ['TAG']
try:
    d=json.loads(open('gpurun_out/%s_bench.json' % t).read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],2), d['e2e']['value'], {k: round(v,2) for k,v in d['stages_ms_last_step'].items()})
    for k in ('sequence258','h36m_1024','train_step','mesh_extract','parity','cpu_baseline'):
        print(k, json.dumps(d.get(k))[:700])
    print('roofline', json.dumps(d['roofline'])[:600])
except Exception as e:
    print('bench parse failed', e)
try:
    d=json.loads(open('gpurun_out/%s_bench_ref.json' % t).read().strip().splitlines()[-1])
    print('ref', d['value'], d['cpu_baseline'])
except Exception as e:
    print('ref parse failed', e)
PY
