# bring-up of the fused decoder tail: layer tests first (bounded), then model tests, then cfg3 / cfg4 A/B
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q --tb=short -x -k "fused_convt" 2>&1 | tail -25 | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_models.py -m gpu -q --tb=short -x -k "fused_decoder or video" 2>&1 | tail -15 | cut -c1-400
P='import json,sys
d=json.loads(sys.stdin.read()); k=d["roofline"]["per_kernel_ms"]
print(d["value"], d["ms_per_step"], d["clocks"]["reasons"], k)'
for v in 0 1; do echo cfg3 FUSE=$v; VAD_FUSE_DEC=$v timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --workload cfg3 2>&1 | tail -1 | python -c "$P"; done
for v in 0 1; do echo cfg4 FUSE=$v; VAD_FUSE_DEC=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --workload cfg4 2>&1 | tail -1 | python -c "$P"; done
