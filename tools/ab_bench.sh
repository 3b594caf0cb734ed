#!/bin/bash
# A/B of whole-step throughput between builds of different commits ON THE SAME BOX (box-to-box clocks differ by +-4 %).
# usage: tools/ab_bench.sh out.txt dir1 dir2 ...   (each dir = a checkout with its own built libavh_b200.so)
out=$1; shift
: > $out
for rep in 1 2; do
  for d in "$@"; do
    extra=""
    grep -q "config3-passes" $d/bench.py && extra="--config3-passes 0"
    grep -q "config5-steps" $d/bench.py && extra="$extra --config5-steps 0"
    grep -q "config4-passes" $d/bench.py && extra="$extra --config4-passes 0"
    line=$(cd $d && python bench.py --no-cpu-baseline --sustain-s 2 --steps 50 --warmup 3 $extra 2>/dev/null | tail -1)
    python - "$d" "$rep" "$line" >> $out <<'PY'
import sys, json
d, rep, line = sys.argv[1:4]
try:
    j = json.loads(line)
    s = j.get("sustained", {})
    print(f"{d:14s} rep{rep} value {j['value']:8.1f} sm {j['clocks']['sm_mhz']} {j['clocks']['reasons']} e2e {j['e2e']['value']:8.1f} "
          f"sustained {s.get('value', 0):8.1f} sm {s.get('clocks', {}).get('sm_mhz')}")
except Exception as e:
    print(d, rep, "ERR", e, line[:200])
PY
  done
done
cat $out
