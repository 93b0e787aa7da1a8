# gpurun --timeout 1200 -- "bash tools/gpu_r2b_1.sh": validate the default build, then mover timing + no-agg experiment + lean wait loops
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_powerlaw.py 128 10 2>&1 | tail -4
i=0
for v in "-DMGN_DEBUG_HOOKS" "-DMGN_WAIT_LEAN" ""; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b1_$i.txt 2>&1
  grep -E "eblk|bwd edge|node|BWD2|FWD3|EPI|MMA|LOADER|MOVER" gpurun_out/r2b1_$i.txt | cut -c1-250
done
