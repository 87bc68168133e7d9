#!/bin/bash
# Closing run of a round: GPU tests, smoke, bench lines (both arms), ncu launch lists and one `--set full`
# capture per hot kernel. The captures are summarised ON the box (tools/ncu_summary.py) so that only the
# summary and one .ncu-rep travel back (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_c2.err
for w in c3 c4; do timeout 600 python bench.py --workload $w --steps 5 --no-extras > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for w in c2 c3; do timeout 300 ncu --metrics $M --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; done
timeout 300 ncu --metrics $M --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
F="--set full --clock-control none"
timeout 600 ncu $F --import-source on -k regex:k_step_tma -s 70 -c 1 -o gpurun_out/prof_c2_fused -f python bench.py --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu $F -k regex:k_step -s 6 -c 1 -o gpurun_out/prof_ms3_bf16 -f python tools/kernel_probe.py --form ms3 --dtype bf16 --reps 4 >> gpurun_out/ncu_bench.log 2>&1
timeout 300 ncu $F -k regex:k_step -s 6 -c 1 -o gpurun_out/prof_ms3_f32 -f python tools/kernel_probe.py --form ms3 --dtype f32 --reps 4 >> gpurun_out/ncu_bench.log 2>&1
for k in k_q_pivots k_q_count k_q_finish; do
  timeout 400 ncu $F -k regex:$k -s 3 -c 1 -o gpurun_out/prof_c4_$k -f python bench.py --workload c4 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
done
python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep --out gpurun_out/ncu_summary > gpurun_out/ncu_summary.stdout 2>&1
find gpurun_out -name '*.ncu-rep' ! -name 'prof_c2_fused.ncu-rep' -delete
du -sh gpurun_out
