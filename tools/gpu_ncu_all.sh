#!/bin/bash
# ncu --set full captures for every kernel on the path (one launch each).  The .ncu-rep files stay on the GPU box (they exceed the
# 64 MiB that travel back); what comes back are the raw metric pages as CSV (gpurun_out/<tag>_*.csv) and the hot source lines of the
# top kernels.  tools/ncu_summary.py csv turns the CSVs into profiles/*.md here.
TAG=${1:-r02n}
mkdir -p gpurun_out /tmp/ncu
COMMON="--no-cpu-baseline --seq-frames 0 --no-h36m"
# A: the render kernels of one steady-state frame (14 arah:: launches per frame; skip the first three frames)
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_trace_|k_iso_|k_knn_|k_corr_|k_shade_|k_sdf_fwd16|k_alpha_cull|k_composite" \
    --launch-skip 42 --launch-count 14 -f -o /tmp/ncu/${TAG}_render \
    python bench.py --steps 2 --warmup 3 --no-train-step --no-mesh $COMMON > gpurun_out/${TAG}_ncu_render.log 2>&1; echo "ncu render rc=$?"
# B: mesh / hypernetwork / ray set-up / image tail
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_sdf_grid|k_mc_|k_hyper_|k_raster_|k_normal_image|k_pose_smpl|k_mask_|k_rays_|k_img_|k_sqdiff|k_ssim_partial|k_project" \
    --launch-count 24 -f -o /tmp/ncu/${TAG}_frows \
    python bench.py --steps 1 --warmup 1 --no-train-step $COMMON > gpurun_out/${TAG}_ncu_frows.log 2>&1; echo "ncu f-rows rc=$?"
# C: training step (one GEMM of each kind is enough) + fused loss
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_gemm_tc|k_loss_partial|k_loss_grads" \
    --launch-skip 600 --launch-count 6 -f -o /tmp/ncu/${TAG}_train \
    python bench.py --steps 1 --warmup 1 --no-mesh $COMMON > gpurun_out/${TAG}_ncu_train.log 2>&1; echo "ncu train rc=$?"
for n in render frows train; do
  ncu -i /tmp/ncu/${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2> /dev/null
  ls -la /tmp/ncu/${TAG}_$n.ncu-rep gpurun_out/${TAG}_${n}_raw.csv
done
# hot source lines (SASS-level sampling, aggregated per CUDA source line) of the three heaviest kernels
for k in k_corr_persist k_trace_persist k_sdf_fwd16 k_shade_tc3; do
  ncu -i /tmp/ncu/${TAG}_render.ncu-rep --page source --csv --kernel-name regex:$k --print-source cuda 2>/dev/null | head -c 3000000 > gpurun_out/${TAG}_src_$k.csv
done
du -sh gpurun_out
