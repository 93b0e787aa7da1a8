N=4
run() { name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-extra "$@" > gpurun_out/bench_${name}_${N}gpu.json 2> gpurun_out/bench_${name}_${N}gpu.err
  tail -c 600 gpurun_out/bench_${name}_${N}gpu.json | head -c 300; echo; grep -i "error\|memory" gpurun_out/bench_${name}_${N}gpu.err | tail -3; }
run c4 --workload c4
run c4_stripes --workload c4 --partition stripes --stripe-rows 25
