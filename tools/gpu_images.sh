#!/bin/bash
# One short GPU-box visit for the rows built after round 1's GPU budget was spent — the image-space tail, the fused
# training loss (§7) and the BASELINE configs[0] fixture: their parity tests, the ncu launch list of the stand-alone image driver and
# one full capture of the rasteriser.  Usage: tools/gpu_images.sh <tag>
TAG=${1:-img}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_zz_images.py -x -q > gpurun_out/${TAG}_pytest_images.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_images.log
tail -5 gpurun_out/${TAG}_pytest_images.log
timeout -s KILL 600 python -m pytest tests/test_gpu_zx_loss.py tests/test_gpu_zw_config0.py -x -q -s > gpurun_out/${TAG}_pytest_loss_config0.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_loss_config0.log
tail -5 gpurun_out/${TAG}_pytest_loss_config0.log
timeout -s KILL 300 python tools/image_tail_profile.py --iters 5 > gpurun_out/${TAG}_image_tail.log 2>&1; echo "driver rc=$?"; tail -5 gpurun_out/${TAG}_image_tail.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_image_launches.csv \
    python tools/image_tail_profile.py --iters 2 > gpurun_out/${TAG}_image_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 300 ncu --set full --import-source on --clock-control none -k regex:k_raster_faces -s 3 -c 1 -f -o gpurun_out/${TAG}_k_raster_faces \
    python tools/image_tail_profile.py --iters 2 > gpurun_out/${TAG}_image_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
