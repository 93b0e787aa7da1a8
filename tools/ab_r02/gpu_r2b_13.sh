# fused sums of the edge backward: one segment at a time (1,4) vs side by side (3,2) -- repeated, same box
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" "-DMGN_BWD2_AGG_AHEAD=3 -DMGN_BWD2_AGG_SIDE=2" "-DMGN_BWD2_AGG_AHEAD=1 -DMGN_BWD2_AGG_SIDE=8"; do
  echo "=== variant '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_PROF_ONLY2=bwd MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 25 2>&1 | grep -E "^bwd edge \(from" | cut -c1-200
done
done
timeout 300 python -m modulus_b200.build > /dev/null
