#!/bin/bash
# Quick check after a kernel change: parity tests, render-only bench line (+ training step), optional extra command.
TAG=${1:-q}
mkdir -p gpurun_out /tmp/ncu
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py -x -q -k "not h36m_1024" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -a "passed\|failed\|Error\|error" gpurun_out/${TAG}_pytest.log | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
TAG=$TAG python - <<'PY'
import json, os
t = os.environ['TAG']
try:
    d=json.loads(open(f'gpurun_out/{t}_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],2), {a: round(v,2) for a,v in d['stages_ms_last_step'].items()})
    print(d['phase_cycles_last_step'].get('trace'))
    ts = d.get('train_step') or {}
    print('train', {k: ts.get(k) for k in ('ms_per_step', 'ms_forward_incl_tracer', 'ms_backward', 'ms_per_step_fused_loss', 'tracer_stages_ms', 'error')})
except Exception as e:
    print('parse failed', e)
PY
tail -3 gpurun_out/${TAG}_bench.err
if [ -n "$2" ]; then
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_hyper_|k_pose_smpl|k_bounds_|k_mask_|k_rays_" --launch-count 24 -f -o /tmp/ncu/${TAG}_setup \
    python bench.py --steps 1 --warmup 1 --no-train-step --no-cpu-baseline --seq-frames 0 --no-h36m > gpurun_out/${TAG}_ncu_setup.log 2>&1; echo "ncu setup rc=$?"
ncu -i /tmp/ncu/${TAG}_setup.ncu-rep --page raw --csv > gpurun_out/${TAG}_setup_raw.csv 2> /dev/null
ls -la gpurun_out/${TAG}_setup_raw.csv
fi
