# edge backward: two helper warps for the fused destination sums (12 segment lanes) vs none -- repeated, same box
mkdir -p gpurun_out
for rep in 1 2; do
for v in "" "-DMGN_BWD2_AGG_HELPERS=0"; do
  echo "=== variant '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_PROF_ONLY2=bwd MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 25 2>&1 | grep -E "^bwd edge \(from" | cut -c1-200
  [ $rep -eq 1 ] && [ -z "$v" ] && timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py tests/test_gpu_fullsize.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -2
done
done
timeout 300 python -m modulus_b200.build > /dev/null
