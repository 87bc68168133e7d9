#!/bin/bash
# Partial closing run after a change to the thresholding kernels: C4 bench line, launch list, quantile captures.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload c4 --steps 5 --no-extras > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 600 python bench.py --workload c4 --steps 40 --no-extras > gpurun_out/bench_c4_sustained.json 2>> gpurun_out/bench_c4.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
for k in k_q_pivots k_q_count k_q_finish; do
  timeout 400 ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -o gpurun_out/prof_c4_$k -f python bench.py --workload c4 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
done
python tools/ncu_summary.py gpurun_out/prof_c4_*.ncu-rep --out gpurun_out/ncu_summary > gpurun_out/ncu_summary.stdout 2>&1
find gpurun_out -name '*.ncu-rep' -delete
