# node block forward on the two-tiles-in-flight kernel (node form) vs the second-generation kernel
mkdir -p gpurun_out
i=0
for v in "" "-DMGN_NODE_FWD2"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  [ $i -eq 1 ] && MGN_NVCC_EXTRA="$v" timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 2>&1 | grep -E "^eblk|^node|^bwd edge \(from" | cut -c1-200
done
timeout 300 python -m modulus_b200.build > /dev/null
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extra > gpurun_out/r2b12_bench.json 2> gpurun_out/r2b12_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2b12_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['whole_step'], d['clocks'], d.get('kernel_shares'))
PY
