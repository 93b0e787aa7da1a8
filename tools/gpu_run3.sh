set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "fused_aggregation or row_ranges" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_kernels.py 1000 1000 3 > gpurun_out/prof_kernels_c3.txt 2>&1; head -12 gpurun_out/prof_kernels_c3.txt
