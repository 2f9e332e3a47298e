#!/usr/bin/env bash
# recurrence-kernel variants: layer time + step trace (scripts/gru_trace.py), GRU parity tests per variant
set -u
PKG=video-based-gait-analysis-for-dementia_b200
for v in "" "$@"; do
  lib=$PKG/lib/libgaitb200${v:+_$v}.so
  echo "== $lib"
  GAITB200_LIB=$lib timeout 120 python scripts/gru_trace.py 2>&1 | grep -E "^rep|^ +[0-9]+ +[-0-9]+ +[-0-9]+ .* [0-9]+$|all CTAs" | sed -n '3,4p;8,9p'
  GAITB200_LIB=$lib timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gru_vs_torch or gru_layer_h0" 2>&1 | tail -1
done
