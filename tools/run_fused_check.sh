# bring-up of the fused decoder tails: layer tests first (bounded), then model tests, then A/B
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --tb=short -x -k "fused_convt" 2>&1 | tail -25 | cut -c1-600
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q --tb=short -x -k "fused_decoder" 2>&1 | tail -15 | cut -c1-400
P='import json,sys
d=json.loads(sys.stdin.read()); k=d["roofline"]["per_kernel_ms"]
print(d["value"], d["ms_per_step"], d["clocks"]["reasons"], k)'
timeout 120 python tools/ablate_tail.py 2>&1 | tail -1
for w in cfg2 cfg3 cfg4; do echo $w; timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload $w 2>&1 | tail -1 | python -c "$P"; done
