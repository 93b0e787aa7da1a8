# gpurun --timeout 1200 -- "bash tools/gpu_r2b_6.sh": fused sums with all bounds fetched ahead + self-publishing gathers, with / without 16-column passes
mkdir -p gpurun_out
i=0
for v in "" "-DMGN_FWD3_PIPE16" "-DMGN_FWD3_SYNC_G" "-DMGN_DEBUG_HOOKS" "-DMGN_DEBUG_HOOKS -DMGN_FWD3_PIPE16"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b6_$i.txt 2>&1
  grep -E "eblk|bwd edge \(from|per-CTA|FWD3|BWD2| EPI   : wMMA2| MMA   : wH1" gpurun_out/r2b6_$i.txt | grep -v "FWD2" | cut -c1-250
  [ $i -le 2 ] && MGN_NVCC_EXTRA="$v" timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
done
timeout 300 python -m modulus_b200.build > /dev/null
