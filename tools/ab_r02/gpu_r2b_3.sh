# gpurun --timeout 1200 -- "bash tools/gpu_r2b_3.sh": A/B of the P_dst prefetch (L1 / L2 / off) and rows in flight of the fused sums
mkdir -p gpurun_out
i=0
for v in "" "-DMGN_FWD3_PF=2" "-DMGN_FWD3_PF=0" "-DMGN_AGG_FLY=4" "-DMGN_DEBUG_HOOKS"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b3_$i.txt 2>&1
  grep -E "eblk|bwd edge \(from|per-CTA|FWD3|BWD2|EPI|MMA   |LOADER" gpurun_out/r2b3_$i.txt | grep -v "FWD2" | cut -c1-250
done
timeout 300 python -m modulus_b200.build > /dev/null
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
