# N-GPU weak-scaling bench: gpurun --gpus N -- "bash tools/gpu_bench_ngpu.sh N"
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/smi8.txt
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err; tail -c 1800 gpurun_out/bench_c3_${N}gpu.json; tail -5 gpurun_out/bench_c3_${N}gpu.err
