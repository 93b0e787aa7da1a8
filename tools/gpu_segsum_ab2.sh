mkdir -p gpurun_out
for v in "-DMGN_LONG_SEG=64" "-DMGN_LONG_SEG=128" "-DMGN_SEG_TAIL=8" "-DMGN_LONG_SEG=64 -DMGN_SEG_TAIL=8"; do
  echo "=== variant '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_powerlaw2.py 2>&1 | grep -E "^all |no hubs|<= 64|uniform 6|hubs only"
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_powerlaw.py 128 10 2>&1 | tail -4
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 2>&1 | grep -E "segsum"
done
timeout 300 python -m modulus_b200.build > /dev/null
