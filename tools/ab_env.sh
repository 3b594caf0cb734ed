#!/bin/bash
# A/B of whole-step throughput between ENVIRONMENT settings of the current build on one box.
# usage: tools/ab_env.sh out.txt "NAME1=VAL ..." "NAME2=VAL ..." ...   ("-" = no extra environment)
out=$1; shift
: > $out
for rep in 1 2; do
  for e in "$@"; do
    envs=""; [ "$e" != "-" ] && envs="$e"
    line=$(env $envs python bench.py --no-cpu-baseline --sustain-s 2 --steps 50 --warmup 3 --config3-passes 0 --config5-steps 0 --config4-passes 0 2>/dev/null | tail -1)
    python - "$e" "$rep" "$line" >> $out <<'PY'
import sys, json
d, rep, line = sys.argv[1:4]
try:
    j = json.loads(line)
    s = j.get("sustained", {})
    print(f"{d:40s} rep{rep} value {j['value']:8.1f} sm {j['clocks']['sm_mhz']} e2e {j['e2e']['value']:8.1f} "
          f"sustained {s.get('value', 0):8.1f} sm {s.get('clocks', {}).get('sm_mhz')}")
except Exception as e:
    print(d, rep, "ERR", e, line[:200])
PY
  done
done
cat $out
