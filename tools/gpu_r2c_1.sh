# (1) Adam comparison diagnostic, fresh processes; (2) the whole GPU tier on the current tree
mkdir -p gpurun_out
for i in 1 2 3 4; do timeout 120 python tools/diag_adam.py 10 2>&1 | tail -12; done > gpurun_out/diag_adam.txt 2>&1
tail -30 gpurun_out/diag_adam.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r2c.log
