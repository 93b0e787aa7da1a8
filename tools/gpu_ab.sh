# A/B of kernel build variants on one box: gpurun --timeout 1200 -- "bash tools/gpu_ab.sh TAG '<nvcc flags A>' '<nvcc flags B>' ..."
# For every flag set: rebuild the library (nvcc is in the image), time the stand-alone kernels at the c3 size (+ per-phase
# cycles when the set contains -DMGN_DEBUG_HOOKS), run the tensor-core tests.  Leaves the default build in place.
tag=$1; shift
mkdir -p gpurun_out
i=0
for v in "$@"; do
  i=$((i+1))
  echo "=== variant $i: '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/ab_${tag}_$i.txt 2>&1
  grep -E "eblk|bwd edge|BWD2|FWD3|EPI|MMA|LOADER" gpurun_out/ab_${tag}_$i.txt | cut -c1-230
  MGN_NVCC_EXTRA="$v" timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -2
done
timeout 300 python -m modulus_b200.build > /dev/null
