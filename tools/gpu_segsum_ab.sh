# gpurun --timeout 1200 -- "bash tools/gpu_segsum_ab.sh": A/B of the segmented-sum forms (default: lane group per segment;
# -DMGN_SEG_OLD: row groups + shuffle tree) on the degree classes of the power-law graph, the C5 shapes and the c3 model sums
mkdir -p gpurun_out
for v in "" "-DMGN_SEG_OLD"; do
  echo "=== variant '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_powerlaw2.py 2>&1 | tail -11
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_powerlaw.py 128 10 2>&1 | tail -4
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 2>&1 | grep -E "segsum"
done
timeout 300 python -m modulus_b200.build > /dev/null
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fused.py tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -4
