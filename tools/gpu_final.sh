#!/usr/bin/env bash
# End-of-round evidence on one GPU box: parity suite, smoke, both bench arms, ncu launch list of the bench command,
# per-kernel metrics of exactly one step, and ONE full ncu capture of the kernel given as $1 (default k_primary_edge).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K=${1:-k_primary_edge}
nvidia-smi -L
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ours.json | cut -c1-300
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.json | cut -c1-400
echo "=== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "=== one profiled step"
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__icc_request_hit_rate.pct,gcc__average_cache_request_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none --csv --log-file gpurun_out/step_kernels.csv python tools/one_step.py > gpurun_out/one_step.log 2>&1
python tools/summarize_step.py gpurun_out/step_kernels.csv gpurun_out/dram_traffic.json | head -16
echo "=== ncu full $K"
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o gpurun_out/prof_$K -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
du -sh gpurun_out
