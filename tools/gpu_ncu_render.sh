#!/bin/bash
# ncu --set full of the render kernels of one steady-state frame only (the A group of tools/gpu_ncu_all.sh)
TAG=${1:-r02z}
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_trace_|k_iso_|k_knn_|k_corr_|k_shade|k_sdf_fwd16|k_alpha_cull|k_composite" --launch-skip 42 --launch-count 14 -f -o /tmp/ncu/${TAG}_render \
    python bench.py --steps 2 --warmup 3 --no-train-step --no-mesh --no-cpu-baseline --seq-frames 0 --no-h36m > gpurun_out/${TAG}_ncu_render.log 2>&1; echo "ncu render rc=$?"
ncu -i /tmp/ncu/${TAG}_render.ncu-rep --page raw --csv > gpurun_out/${TAG}_render_raw.csv 2> /dev/null
ls -la gpurun_out/${TAG}_render_raw.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-train-step --no-cpu-baseline --no-mesh --seq-frames 0 --no-h36m > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
