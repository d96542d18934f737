#!/bin/bash
# A/B of ranking-kernel switches on one box: alternating processes; usage: rank_ab_env.sh "VAR=a" "VAR=b" ...
for i in 1 2 3; do
  for kv in "$@"; do
    env $kv PROBE_Q3=${PROBE_Q3:-100000} PROBE_ONLY=0:4 PROBE_ENV="$kv" python tools/rank_probe.py 2>&1 | grep "^pair"
  done
done
