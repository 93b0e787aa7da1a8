# A/B of the mbarrier wait-loop variants (profiles/r01_issue_slots.md): gpurun --timeout 900 -- "bash tools/gpu_ab_wait.sh"
# Rebuilds the library on the box (nvcc is in the image, ~20 s per build), times the stand-alone kernels at the c3 size
# with each variant and leaves the default build in place.  Compare the "bwd edge (from h1)" / "eblk fwd3" lines.
set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py 1000 1000 3 > gpurun_out/ab_wait_default.txt 2>&1; head -8 gpurun_out/ab_wait_default.txt
for v in "-DMGN_WAIT_HINT=20000" "-DMGN_WAIT_SLEEP=32" "-DMGN_WAIT_SLEEP=200"; do
  tag=$(echo "$v" | tr -c 'A-Za-z0-9\n' '_')
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build || continue
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_kernels.py 1000 1000 3 > gpurun_out/ab_wait$tag.txt 2>&1; head -8 gpurun_out/ab_wait$tag.txt
  MGN_NVCC_EXTRA="$v" timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -1
done
timeout 300 python -m modulus_b200.build   # default flags again
