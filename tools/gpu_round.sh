#!/bin/bash
# Round-2 GPU pass: tools/gpu_round.sh TAG [what...]   what = tests bench prof sanitize (default: all)
TAG=$1; shift
WHAT=${@:-tests bench prof sanitize}
mkdir -p gpurun_out
for w in $WHAT; do
case $w in
tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/tests_$TAG.log;;
bench) timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_$TAG.log;;
prof) bash tools/profile_round.sh $TAG;;
sanitize)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck_$TAG.log
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck_$TAG.log;;
esac
done
