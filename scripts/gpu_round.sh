#!/bin/bash
# One gpurun session: parity tests, kernel tests, bench, launch list.  Every leg has its own timeout and log.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout -s KILL 1200 python -m pytest tests -q -m gpu ${PYTEST_ARGS:-} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
grep -E "passed|failed|Error|error|assert" gpurun_out/pytest.log | tail -15
if [ "${BENCH:-1}" = "1" ]; then
echo "== bench" ; timeout -s KILL 900 python bench.py --steps ${BENCH_STEPS:-2} --warmup 3 --precision ${PRECISION:-bf16} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
fi
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-400} --csv --log-file gpurun_out/launches.csv \
     python scripts/ncu_target.py > gpurun_out/ncu_target.log 2>&1; echo "ncu rc=$?"
  tail -3 gpurun_out/ncu_target.log
fi
if [ "${NCU_FULL:-0}" = "1" ]; then
  echo "== ncu full (top kernels)"
  timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
     -k regex:'rnn_tc_kernel|gemm_tc2?_kernel|conv_tc_kernel|spectrogram_kernel' -c ${NCU_FULL_COUNT:-8} -f -o gpurun_out/prof_top \
     python scripts/ncu_target.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
  tail -2 gpurun_out/ncu_full.log
fi
if [ "${SMOKE:-0}" = "1" ]; then
  echo "== smoke"; timeout -s KILL 300 python __graft_entry__.py smoke 2>&1 | tail -4
fi
