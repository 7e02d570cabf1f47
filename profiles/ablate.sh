run() { python bench.py --engine ${ENGINE:-auto} --candidates 2097152 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$1', 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; }
BX_TC_ABLATE=0 run full; BX_TC_ABLATE=1 run no_kstar_math; BX_TC_ABLATE=2 run no_mma; BX_TC_ABLATE=3 run neither
