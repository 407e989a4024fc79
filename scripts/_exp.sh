timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "groups_in_flight or merged or carry_state or large_batch" 2>&1 | tail -4
export LAYERS=3 REPS=2 BATCHES=256 NIFS=2 KSPLITS=2
for cfg in "0" "1" "7"; do
  echo "== skip=$cfg"
  DSB_RNN_SKIP=$cfg timeout 100 python scripts/rnn_ab.py 2>&1 | grep rnn_ms
done
BATCHES=384 NIFS=3 timeout 100 python scripts/rnn_ab.py 2>&1 | grep rnn_ms
DSB_RNN_DEBUG=1 LAYERS=1 REPS=1 timeout 100 python scripts/rnn_ab.py 2>&1 | tail -28
