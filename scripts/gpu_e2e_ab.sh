#!/usr/bin/env bash
# A/B of the end-to-end number on ONE box: alternate library builds / PDL masks, print e2e and device-timed values
set -u
PKG=video-based-gait-analysis-for-dementia_b200
for rep in 1 2; do
for cfg in ":1" "_acopy0:1" ":0" "_acopy0:0"; do
  v=${cfg%%:*}; p=${cfg#*:}
  GAITB200_PDL=$p GAITB200_LIB=$PKG/lib/libgaitb200$v.so timeout 300 python bench.py --steps 30 --warmup 5 --no-extra-configs --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('lib=${v:-default} pdl=$p value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(e['value']), 'd2h', round(e['d2h_gbs'],1), 'ceil', round(e['d2h_ceiling_gbs'],1))"
done
done
