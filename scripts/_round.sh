set -u
export DSB_RNN_NONCOOP=1 PRECISION=bf16 NCU_BATCH=256
echo "== ncu launch list (256 sequences per pass)"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python scripts/ncu_target.py > gpurun_out/ncu_target_r2b.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_target_r2b.log
python scripts/launch_summary.py gpurun_out/launches_r2b.csv
echo "== ncu full: pair kernel, rows gemm, combine"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:'rnn_pair_kernel|gemm_tc2_kernel|combine_dirs_t' -c 6 -f -o gpurun_out/prof_pair python scripts/ncu_target.py > gpurun_out/ncu_full_r2b.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full_r2b.log
ls -la gpurun_out/prof_pair.ncu-rep
