# gpurun --timeout 1200 -- "bash tools/gpu_r2b_7.sh": h1 store warps in the edge forward, with / without 16-column passes; bounds-ahead depth of the fused sums
mkdir -p gpurun_out
i=0
for v in "" "-DMGN_FWD3_PIPE16" "-DMGN_FWD3_PIPE16 -DMGN_FWD3_MOVER_H1" "-DMGN_SEG_AHEAD=1" "-DMGN_DEBUG_HOOKS -DMGN_FWD3_PIPE16"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b7_$i.txt 2>&1
  grep -E "eblk|bwd edge \(from|per-CTA|FWD3" gpurun_out/r2b7_$i.txt | grep -v "FWD2" | cut -c1-250
  [ $i -le 2 ] && MGN_NVCC_EXTRA="$v" timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
done
timeout 300 python -m modulus_b200.build > /dev/null
