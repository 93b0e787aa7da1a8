# 2-GPU check of the partitioned path: gpurun --gpus 2 -- 'bash tools/gpu_dist_check.sh'
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/pytest_gpu_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_dist.log
tail -5 gpurun_out/pytest_gpu_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; tail -c 700 gpurun_out/bench_c3_2gpu.json
