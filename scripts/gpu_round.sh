#!/bin/bash
# One gpurun session: parity tests, bench, launch list.  Every leg has its own timeout and log.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest" ; timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -15 gpurun_out/pytest.log
echo "== bench" ; timeout 600 python bench.py --steps ${BENCH_STEPS:-2} --warmup 3 --precision ${PRECISION:-fp32} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ "${NCU:-1}" = "1" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-400} --csv --log-file gpurun_out/launches.csv \
     python scripts/ncu_target.py > gpurun_out/ncu_target.log 2>&1; echo "ncu rc=$?"
  tail -3 gpurun_out/ncu_target.log
fi
