#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, one full ncu capture.  Usage: tools/gpu_check.sh <tag> [kernel-regex]
TAG=${1:-run}
KRE=${2:-k_corr_tc4}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
head -c 1500 gpurun_out/${TAG}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-train-step --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:${KRE} -s 1 -c 1 -f -o gpurun_out/${TAG}_${KRE} \
    python bench.py --steps 1 --warmup 1 --no-train-step --no-cpu-baseline --no-mesh > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -12
