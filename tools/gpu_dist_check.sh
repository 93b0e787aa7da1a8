# Multi-GPU check of the partitioned path with real NCCL: gpurun --gpus N -- 'bash tools/gpu_dist_check.sh N'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
[ -z "$SKIP_TESTS" ] && timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_dist_${N}gpu_nccl.log
run() {  # name, bench args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-extra "$@" > gpurun_out/bench_${name}_${N}gpu.json 2> gpurun_out/bench_${name}_${N}gpu.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_${name}_${N}gpu.json") if l.startswith("{")][-1]
    c=d["config"]
    print("${name} N=${N}:", round(d["value"]/1e6,1), "M edges/s", round(d["ms_per_step"],1), "ms", d["scaling"], "| partition:", c.get("partition"), "| halo max frac", round(d.get("halo",{}).get("fraction_max",0),4), "| e2e", round(d["e2e"]["value"]/1e6,1))
except Exception as e:
    print("${name} N=${N}: FAILED", e); print(open("gpurun_out/bench_${name}_${N}gpu.err").read()[-1500:])
PY
}
run c3 --workload c3
run c3_stripes --workload c3 --partition stripes --stripe-rows 25
if [ "$N" -ge 4 ]; then
  run c4 --workload c4
  run c4_stripes --workload c4 --partition stripes --stripe-rows 25
fi
