#!/bin/bash
# Tests + bench lines + ncu launch lists (no full captures).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_c2.err
for w in c3 c4; do timeout 600 python bench.py --workload $w --steps 5 --no-extras > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for w in c2 c3; do timeout 300 ncu --metrics $M --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; done
timeout 300 ncu --metrics $M --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
