#!/usr/bin/env bash
# recurrence-kernel variants: layer time + step trace (scripts/gru_trace.py)
set -u
PKG=video-based-gait-analysis-for-dementia_b200
for v in "" "$@"; do
  lib=$PKG/lib/libgaitb200${v:+_$v}.so
  echo "== $lib"
  GAITB200_LIB=$lib timeout 120 python scripts/gru_trace.py 2>&1 | grep -E "^rep|^ +[0-9]+ +[-0-9]+ +[-0-9]+ .* [0-9]+$|all CTAs" | sed -n '2,4p;8,10p;20,21p'
done
