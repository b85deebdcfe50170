#!/bin/bash
# compute-sanitizer memcheck over a few small GPU tests (out-of-bounds / misaligned accesses in the kernels of the render path and the banded lattice)
TAG=${1:-san}
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 3 \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_mesh.py -x -q -k "(golden and zju377 and tf32) or (banded and 17) or per_step or (mixed and tf32-fp32)" \
    > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.log
grep -a "ERROR SUMMARY\|passed\|failed\|Invalid\|out of bounds\|misaligned" gpurun_out/${TAG}_memcheck.log | head -20
