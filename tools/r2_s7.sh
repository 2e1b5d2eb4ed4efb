set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/trace_once.py bistro 1024 cwbvh > gpurun_out/t_bistro_cwbvh_bf16.txt 2>&1; cat gpurun_out/t_bistro_cwbvh_bf16.txt
timeout 300 python tools/quick_build.py sponza bunny bistro > gpurun_out/build0.log 2>&1; tail -6 gpurun_out/build0.log
TBVH_BUILD_MODE=1 timeout 300 python tools/quick_build.py sponza bunny bistro > gpurun_out/build1.log 2>&1; tail -6 gpurun_out/build1.log
