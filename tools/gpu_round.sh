#!/bin/bash
# One GPU-box visit: parity tests, smoke, the bench with default flags, the reference arm, optional ncu launch list / full capture.  Usage: tools/gpu_round.sh TAG [ref] [ncu] [ncufull] [ncuk0] [load]
TAG=${1:-x}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests_$TAG.log; tail -3 gpurun_out/tests_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_$TAG.json; echo; tail -3 gpurun_out/bench_$TAG.err
for a in "$@"; do
  case $a in
    ref)
      timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$?"; head -c 700 gpurun_out/bench_ref_$TAG.json; echo;;
    ncu)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs > gpurun_out/ncu_l_$TAG.log 2>&1; echo "ncu list rc=$?";;
    ncufull)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k1_score_kernel|k1_ladder_kernel' -s 2 -c 2 -f -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs > gpurun_out/ncu_f_$TAG.log 2>&1; echo "ncu full rc=$?";;
    ncuk0)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k0_prepare_kernel' -s 4 -c 2 -f -o gpurun_out/prof_k0_$TAG python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 --no-configs > gpurun_out/ncu_k0_$TAG.log 2>&1; echo "ncu k0 rc=$?";;
    load)
      timeout 600 python tools/prof_e2e2.py > gpurun_out/load_$TAG.log 2>&1; echo "load rc=$?"; tail -4 gpurun_out/load_$TAG.log;;
  esac
done
