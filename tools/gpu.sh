#!/usr/bin/env bash
# One parameterised GPU-box driver (replaces the per-call scripts of round 1). Run through gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu.sh tests "-k real_shapes"'
#   gpurun --timeout 900  -- 'bash tools/gpu.sh bench "--steps 20 --warmup 5"'
#   gpurun --timeout 900  -- 'bash tools/gpu.sh ncu_launches'         # launch list of one PC step
#   gpurun --timeout 900  -- 'bash tools/gpu.sh ncu_full <kernel-regex> <target.py>'
# Everything it writes goes to gpurun_out/ (merged back by gpurun); summaries worth keeping are copied to profiles/.
set -u
mkdir -p gpurun_out
what=${1:-tests}; shift || true
case "$what" in
  tests)   python -m pytest tests -m gpu -x -q ${1:-} 2>&1 | tee gpurun_out/pytest_gpu.log | tail -40 ;;
  bench)   python bench.py ${1:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err ;;
  smoke)   python __graft_entry__.py smoke 2>&1 | tail -5 ;;
  ncu_launches)
           ncu --metrics gpu__time_duration.sum --clock-control none -c ${1:-1500} --csv --log-file gpurun_out/launches.csv \
               python tools/step_ncu_target.py > gpurun_out/ncu_launches.log 2>&1
           python tools/summarize_ncu_csv.py gpurun_out/launches.csv | tee gpurun_out/launches.md | tail -40 ;;
  ncu_full)
           ncu --set full --clock-control none --import-source on -k "regex:${1}" -c ${3:-3} -o gpurun_out/full_${1} -f \
               python ${2} > gpurun_out/ncu_full.log 2>&1; ls -la gpurun_out | tail -5 ;;
  sanitizer)
           compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -x -q ${1:-} \
               > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/sanitizer.log ;;
  *)       echo "unknown mode $what"; exit 2 ;;
esac
