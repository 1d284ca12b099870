P='import json,sys
d=json.loads(sys.stdin.read()); k=d["roofline"]["per_kernel_ms"]
print(d["value"], d["ms_per_step"], d["clocks"]["reasons"], k)'
for w in cfg2 cfg3 cfg4 cfg2; do echo $w; timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python -c "$P"; done
