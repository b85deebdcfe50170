#!/bin/bash
# compute-sanitizer memcheck over a few small GPU tests (out-of-bounds / misaligned accesses in the kernels of the render path and the banded lattice)
TAG=${1:-san}
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 3 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py tests/test_gpu_train.py tests/test_gpu_zz_images.py -x -q -k "not 512 and not 1024 and not full_size and not 256 and not 200000 and not 20000" \
    > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.log
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds\|misaligned" gpurun_out/${TAG}_memcheck.log | head -20
