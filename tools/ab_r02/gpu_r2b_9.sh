# per-phase cycles of a SLOW CTA (blockIdx 100) of the edge forward
mkdir -p gpurun_out
for v in "-DMGN_DEBUG_HOOKS -DMGN_FWD3_PIPE16" "-DMGN_DEBUG_HOOKS -DMGN_FWD3_PIPE16 -DMGN_FWD3_MOVER_H1" "-DMGN_DEBUG_HOOKS"; do
  for cta in 0 100; do
  echo "=== variant '$v' timing CTA $cta"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_TIMING_CTA=$cta MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 5 2>&1 | grep -E "^eblk|FWD3" | cut -c1-250
  done
done
timeout 300 python -m modulus_b200.build > /dev/null
