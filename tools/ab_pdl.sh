timeout 600 python -m pytest tests -m gpu -q --tb=line -x 2>&1 | tail -2 | cut -c1-300
P='import json,sys
d=json.loads(sys.stdin.read()); k=d["roofline"]["per_kernel_ms"]
print(d["value"], d["ms_per_step"], d["clocks"]["reasons"], k)'
for v in 0 1 0 1; do echo PDL_ALL=$v; VAD_PDL_ALL=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$P"; done
for v in 0 1; do echo cfg3 PDL_ALL=$v; VAD_PDL_ALL=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | python -c "$P"; done
for v in 0 1; do echo cfg4 PDL_ALL=$v; VAD_PDL_ALL=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg4 2>&1 | tail -1 | python -c "$P"; done
