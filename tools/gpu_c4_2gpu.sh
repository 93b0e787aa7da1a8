# c4 (one 8 M-node / 48 M-edge mesh) on 2 GPUs: gpurun --gpus 2 --timeout 1500 -- "bash tools/gpu_c4_2gpu.sh"   (memory-lean mode: 24 M edges per rank; expandable allocator segments: the step peaks at ~165 of 178 GiB)
N=2
mkdir -p gpurun_out
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-extra --no-cpu --workload c4 > gpurun_out/bench_c4_${N}gpu.json 2> gpurun_out/bench_c4_${N}gpu.err
tail -c 2500 gpurun_out/bench_c4_${N}gpu.json; echo; grep -i "error\|memory" gpurun_out/bench_c4_${N}gpu.err | tail -3
nvidia-smi --query-gpu=memory.used --format=csv | tail -2
