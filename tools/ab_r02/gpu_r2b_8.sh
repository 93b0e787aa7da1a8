# per-CTA (per-SM) cycles per tile of the edge forward with fused sums
mkdir -p gpurun_out
for v in "-DMGN_DEBUG_HOOKS -DMGN_FWD3_PIPE16 -DMGN_FWD3_MOVER_H1" "-DMGN_DEBUG_HOOKS"; do
  echo "=== variant '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_PROF_DUMP=1 MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 5 2>&1 | grep -E "by SM id|per-CTA" | cut -c1-2500
done
timeout 300 python -m modulus_b200.build > /dev/null
