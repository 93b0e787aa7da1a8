# compute-sanitizer memcheck over the GPU test tier (small sizes; the full-size and multi-process files are left out):
#   gpurun --timeout 600 -- 'bash tools/gpu_sanitizer.sh'
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_smoke.log
tail -3 gpurun_out/sanitizer_smoke.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_fused.py tests/test_gpu_tc.py tests/test_gpu_ops.py tests/test_train_step.py tests/test_graphcast.py tests/test_rollout.py tests/test_halo_partition.py -m gpu -q > gpurun_out/sanitizer_tests.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_tests.log
tail -8 gpurun_out/sanitizer_tests.log
