# gpurun --timeout 1200 -- "bash tools/gpu_r2b_4.sh": early accumulator-drain signal in the edge backward, two-stage index pipeline of the fused sums
mkdir -p gpurun_out
i=0
for v in "" "-DMGN_DEBUG_HOOKS"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b4_$i.txt 2>&1
  grep -E "eblk|bwd edge \(from|node|per-CTA|FWD3|BWD2|EPI|MMA   |LOADER" gpurun_out/r2b4_$i.txt | grep -v "FWD2" | cut -c1-250
done
timeout 300 python -m modulus_b200.build > /dev/null
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_ops.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b4_bench.json 2> gpurun_out/r2b4_bench.err; tail -c 1500 gpurun_out/r2b4_bench.json
