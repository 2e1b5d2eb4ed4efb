set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cwbvh_gpu.py tests/test_variants_gpu.py tests/test_build_gpu.py tests/test_trace_gpu.py tests/test_bistro_gpu.py tests/test_group_gpu.py -m gpu -q -x > gpurun_out/pytest_gpu9.log 2>&1; tail -12 gpurun_out/pytest_gpu9.log
timeout 300 python tools/trace_once.py bistro 1024 cwbvh --stats > gpurun_out/t_bistro_cwbvh_v3.txt 2>&1; cat gpurun_out/t_bistro_cwbvh_v3.txt
timeout 300 python tools/trace_once.py sponza 1024 cwbvh --stats > gpurun_out/t_sponza_cwbvh_v3.txt 2>&1; cat gpurun_out/t_sponza_cwbvh_v3.txt
timeout 300 python tools/quick_build.py sponza bunny bistro lucy_dragon_x29 > gpurun_out/build_hybrid.log 2>&1; tail -4 gpurun_out/build_hybrid.log
