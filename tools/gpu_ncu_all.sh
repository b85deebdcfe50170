#!/bin/bash
# ncu --set full captures for every kernel on the path (one launch each).  The .ncu-rep files stay on the GPU box (they exceed the
# 64 MiB that travel back); what comes back are the raw metric pages as CSV (gpurun_out/<tag>_*.csv).
# tools/ncu_summary.py csv turns the CSVs into profiles/*.md (and profiles/r02_traffic.json) here.
TAG=${1:-r02p}
mkdir -p gpurun_out /tmp/ncu
COMMON="--no-cpu-baseline --seq-frames 0 --no-h36m"
cap() {   # cap <name> <kernel regex> <skip> <count> <bench flags...>
  local name=$1 re=$2 skip=$3 cnt=$4; shift 4
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:"$re" --launch-skip $skip --launch-count $cnt -f -o /tmp/ncu/${TAG}_$name \
      python bench.py "$@" $COMMON > gpurun_out/${TAG}_ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i /tmp/ncu/${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2> /dev/null
  ls -la /tmp/ncu/${TAG}_$name.ncu-rep gpurun_out/${TAG}_${name}_raw.csv
}
# A: the render kernels of one steady-state frame (k_knn_build + 13 render launches per frame; skip the first three frames)
cap render "k_trace_|k_iso_|k_knn_|k_corr_|k_shade|k_sdf_fwd16|k_alpha_cull|k_composite" 42 14 --steps 2 --warmup 3 --no-train-step --no-mesh
# B: the rows either side of the renderer, one group per capture so that repeated launches of one kernel cannot use up the count
cap setup "k_hyper_|k_pose_smpl|k_mask_|k_rays_|k_project|k_pack_f16|k_layer_scale|k_pack_film" 0 24 --steps 1 --warmup 1 --no-train-step
cap mesh "k_sdf_grid16|k_grid_band|k_mc_" 0 16 --steps 1 --warmup 1 --no-train-step
cap image "k_raster_|k_normal_image|k_img_|k_sqdiff|k_ssim" 0 16 --steps 1 --warmup 1 --no-train-step
# C: training step (one GEMM of each kind is enough) + fused loss
cap train "k_gemm_tc" 600 4 --steps 1 --warmup 1 --no-mesh
cap loss "k_loss_partial|k_loss_grads|k_loss_" 0 4 --steps 1 --warmup 1 --no-mesh
du -sh gpurun_out
