# memory-lean mode: single-device and partitioned tests (ranks share one GPU)
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -5
