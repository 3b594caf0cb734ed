for s in 1 2 3 4; do python bench.py --no-cpu-baseline --sustain-s 2 --steps 50 --warmup 3 --config3-passes 0 --config5-steps 0 --streams $s 2>/dev/null | tail -1 | python -c "
import sys,json; j=json.loads(sys.stdin.read()); print('streams $s value %.1f e2e %.1f sustained %.1f sm %s' % (j['value'], j['e2e']['value'], j['sustained']['value'], j['sustained']['clocks']['sm_mhz']))"; done
