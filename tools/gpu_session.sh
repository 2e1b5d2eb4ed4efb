set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_build.py sponza bistro 2>&1 | grep -E "build"
