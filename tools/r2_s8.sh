set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants_gpu.py tests/test_build_gpu.py tests/test_bistro_gpu.py -m gpu -q -x -s > gpurun_out/pytest_gpu8.log 2>&1; tail -12 gpurun_out/pytest_gpu8.log
for c in 1 2 4 8; do TBVH_BUILD_CTAS=$c timeout 300 python tools/quick_build.py sponza bunny bistro lucy_dragon_x29 > gpurun_out/build_c$c.log 2>&1; echo "ctas $c"; tail -4 gpurun_out/build_c$c.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 2 -c 1 -o gpurun_out/r2_cw_primary python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets primary > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 3 -c 1 -o gpurun_out/r2_cw_shadow python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets primary,shadow > /dev/null 2>&1
ls -la gpurun_out | tail -8
