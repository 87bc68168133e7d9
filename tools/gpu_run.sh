#!/bin/bash
# One script for every GPU call of a round (run under gpurun). Stages are picked by name:
#   bash tools/gpu_run.sh tests smoke bench ref launches ncu_step ncu_quantile sweep latency
# Outputs go to gpurun_out/ (merged back by gpurun); ncu captures are summarised ON the box
# (tools/ncu_summary.py) so only the summaries and one .ncu-rep travel back (64 MiB cap).
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
F="--set full --clock-control none"
for stage in "$@"; do
case $stage in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log ;;
tests_all)
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -30 ;;
smoke)
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
  python tools/bench_digest.py gpurun_out/bench_default.json ;;
bench_each)
  for w in c2 c3 c4; do timeout 600 python bench.py --workload $w --steps 5 --no-extras > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python tools/bench_digest.py gpurun_out/bench_$w.json; done ;;
ref)
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; python tools/bench_digest.py gpurun_out/bench_ref.json ;;
launches)
  for w in c2 c3; do timeout 300 ncu --metrics $M --clock-control none -s 60 -c 44 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; done
  timeout 300 ncu --metrics $M --clock-control none -s 300 -c 100 --csv --log-file gpurun_out/launches_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1 ;;
ncu_step)
  timeout 600 ncu $F --import-source on -k regex:k_step_tma -s 70 -c 1 -o gpurun_out/prof_c2_fused -f python bench.py --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
  timeout 600 ncu $F --import-source on -k regex:k_step_tma -s 50 -c 1 -o gpurun_out/prof_c3_fused -f python bench.py --workload c3 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
  timeout 300 ncu $F -k regex:k_step -s 6 -c 1 -o gpurun_out/prof_ms3_bf16 -f python tools/kernel_probe.py --form ms3 --dtype bf16 --reps 4 >> gpurun_out/ncu_bench.log 2>&1
  timeout 300 ncu $F -k regex:k_step -s 6 -c 1 -o gpurun_out/prof_ms3_f32 -f python tools/kernel_probe.py --form ms3 --dtype f32 --reps 4 >> gpurun_out/ncu_bench.log 2>&1 ;;
ncu_quantile)
  for k in k_q_pivots k_q_count k_q_finish k_thr; do
    timeout 400 ncu $F -k regex:$k -s 3 -c 1 -o gpurun_out/prof_c4_$k -f python bench.py --workload c4 --steps 1 --warmup 3 --no-extras >> gpurun_out/ncu_bench.log 2>&1
  done ;;
sweep)
  timeout 1200 bash tools/inloop_sweep.sh c2 c3 > /dev/null 2>&1; cat gpurun_out/inloop_sweep.txt ;;
l2ncu)
  timeout 400 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:k_step_direct --csv --log-file gpurun_out/l2_ncu.csv python tools/l2_group_probe.py 16 24 32 48 64 > gpurun_out/l2_ncu.log 2>&1; python tools/l2_ncu_digest.py gpurun_out/l2_ncu.csv ;;
l2probe)
  timeout 300 python tools/l2_group_probe.py > gpurun_out/l2_group_probe.txt 2>&1; cat gpurun_out/l2_group_probe.txt ;;
adaptive_probe)
  timeout 600 python tools/adaptive_probe.py > gpurun_out/adaptive_probe.txt 2>&1; cat gpurun_out/adaptive_probe.txt ;;
sanitize)
  for t in memcheck racecheck synccheck; do timeout 900 compute-sanitizer --tool $t python tools/sanitize_probe.py > gpurun_out/sanitizer_$t.log 2>&1; echo "$t rc=$?" >> gpurun_out/sanitizer_$t.log; tail -3 gpurun_out/sanitizer_$t.log; done ;;
latency)
  timeout 600 python tools/host_overhead.py > gpurun_out/host_overhead.txt 2>&1; cat gpurun_out/host_overhead.txt ;;
*) echo "unknown stage $stage" ;;
esac
done
if ls gpurun_out/prof_*.ncu-rep > /dev/null 2>&1; then
  python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep --out gpurun_out/ncu_summary > gpurun_out/ncu_summary.stdout 2>&1
  find gpurun_out -name '*.ncu-rep' ! -name 'prof_c2_fused.ncu-rep' -delete
fi
du -sh gpurun_out
