#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, optional sanitizer pass / ncu launch list.  Usage: tools/gpu_round.sh TAG [sanitize] [ncu]
TAG=${1:-x}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests_$TAG.log; tail -3 gpurun_out/tests_$TAG.log
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 1500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for a in "$@"; do
  case $a in
    sanitize)
      for tool in memcheck racecheck; do
        timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/sanitizer_${tool}_$TAG.log 2>&1; echo "$tool rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_$TAG.log
      done;;
    ncu)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list rc=$?";;
    ncufull)
      ncu --set full --clock-control none --import-source on -k regex:k1_score_kernel -s 4 -c 2 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu full rc=$?";;
    load)
      python tools/prof_load2.py > gpurun_out/load_$TAG.log 2>&1; echo "load rc=$?"; cat gpurun_out/load_$TAG.log | tail -12;;
  esac
done
