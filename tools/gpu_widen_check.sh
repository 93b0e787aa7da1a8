# New-row check: gpurun --timeout 420 -- "bash tools/gpu_widen_check.sh"
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_graphcast.py tests/test_train_step.py -m gpu -q > gpurun_out/pytest_widen.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_widen.log
tail -40 gpurun_out/pytest_widen.log
timeout 150 python tools/bench_train_step.py > gpurun_out/train_step.md 2> gpurun_out/train_step.err; cat gpurun_out/train_step.md; tail -5 gpurun_out/train_step.err
