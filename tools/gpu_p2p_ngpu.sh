# gpurun --gpus N --timeout 900 -- "bash tools/gpu_p2p_ngpu.sh N": peer-memory halo exchange on 2 GPUs -- parity test, then c3 stripes bench NCCL vs P2P
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q -k "peer_memory" 2>&1 | tail -5
N=${1:-2}
for mode in 0 1; do
  MGN_HALO_P2P=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode bench.py --gpus $N --steps 5 --warmup 3 --no-extra --no-cpu --partition stripes --stripe-rows 25 > gpurun_out/bench_c3_stripes_${N}gpu_p2p$mode.json 2> gpurun_out/bench_c3_stripes_${N}gpu_p2p$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c3_stripes_${N}gpu_p2p$mode.json').read().strip().splitlines()[-1])
    print("P2P=$mode", d['value'], d['ms_per_step'], d.get('halo',{}).get('fraction_max'), d['gpu_launches'])
except Exception as e:
    print("P2P=$mode failed", e); print(open('gpurun_out/bench_c3_stripes_${N}gpu_p2p$mode.err').read()[-1500:])
PY
done
