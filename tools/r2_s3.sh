set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_variants_gpu.py tests/test_tlas_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu3.log 2>&1; tail -15 gpurun_out/pytest_gpu3.log
timeout 300 python tools/pcie_probe2.py 24 > gpurun_out/pcie2.txt 2>&1; cat gpurun_out/pcie2.txt
