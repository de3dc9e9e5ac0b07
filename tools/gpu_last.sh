#!/usr/bin/env bash
# one short box session: the GPU suite (captured prints of the passing tests included), the C3/C4 goldens once more with the host-built
# edge list, and two bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -q -rP > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log | cut -c1-300
grep -h "edge build ms\|FAILED\|Error" gpurun_out/r02_pytest_final.log | cut -c1-260 | head -30
echo "=== C3/C4 goldens, host edge list"; RB_HOST_EDGE_LIST=1 timeout 200 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "c3_ or c4_ or bvh_stress" 2>&1 | tail -2 | cut -c1-200
echo "=== bench c3"; timeout 200 python bench.py --steps 1 --warmup 3 --workload c3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c3_device_list.json | cut -c1-1200
echo "=== bench c2"; timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours_n1_c2_final.json | cut -c1-700
