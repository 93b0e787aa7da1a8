# gpurun --timeout 1200 -- "bash tools/gpu_r2b_5.sh": 16-column software-pipelined TMEM passes revisited (time AND cycles AND SM clock per variant)
mkdir -p gpurun_out
i=0
for v in "" "-DMGN_FWD3_PIPE16" "-DMGN_BWD2_PIPE16" "-DMGN_DEBUG_HOOKS" "-DMGN_DEBUG_HOOKS -DMGN_FWD3_PIPE16" "-DMGN_DEBUG_HOOKS -DMGN_BWD2_PIPE16"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b5_$i.txt 2>&1
  grep -E "eblk|bwd edge \(from|per-CTA|FWD3|BWD2| EPI   : wMMA2| MMA   : wH1" gpurun_out/r2b5_$i.txt | grep -v "FWD2" | cut -c1-250
done
timeout 300 python -m modulus_b200.build > /dev/null
