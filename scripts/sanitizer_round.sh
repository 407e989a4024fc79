#!/bin/bash
# compute-sanitizer pass over the round's kernels (run under gpurun; logs to gpurun_out/sanitizer_*.log)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 600 compute-sanitizer --tool $tool python __graft_entry__.py smoke > gpurun_out/sanitizer_smoke_$tool.log 2>&1
  echo "== $tool smoke"; grep -E "SUMMARY|smoke\] ok" gpurun_out/sanitizer_smoke_$tool.log | tail -3
done
# the fp32 step kernel, both persistent recurrences with groups in flight, the .klm path, streaming state carry
SEL='headline_fp32 or (in_flight and gru-150) or klm_equals_arpa or streaming_bf16_groups or windows'
timeout -s KILL 1500 compute-sanitizer --tool memcheck python -m pytest tests -q -m gpu -k "$SEL" > gpurun_out/sanitizer_tests_memcheck.log 2>&1
echo "== memcheck tests"; grep -E "passed|failed|SUMMARY" gpurun_out/sanitizer_tests_memcheck.log | tail -4
SEL2='(in_flight and gru-150) or streaming_bf16_groups'
timeout -s KILL 1200 compute-sanitizer --tool racecheck python -m pytest tests -q -m gpu -k "$SEL2" > gpurun_out/sanitizer_tests_racecheck.log 2>&1
echo "== racecheck tests"; grep -E "passed|failed|SUMMARY" gpurun_out/sanitizer_tests_racecheck.log | tail -4
