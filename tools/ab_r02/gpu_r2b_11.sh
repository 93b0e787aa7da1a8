mkdir -p gpurun_out
i=0
for v in "-DMGN_FWD3_PIPE16" "-DMGN_FWD3_PIPE16 -DMGN_FWD3_MOVER_H1" "-DMGN_FWD3_PIPE16 -DMGN_FWD3_IDLE_STORE" "-DMGN_FWD3_PIPE16" "-DMGN_FWD3_PIPE16 -DMGN_FWD3_MOVER_H1" "-DMGN_FWD3_PIPE16 -DMGN_FWD3_IDLE_STORE"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_PROF_ONLY2=1 MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 25 2>&1 | grep -E "^eblk" | cut -c1-250
done
timeout 300 python -m modulus_b200.build > /dev/null
