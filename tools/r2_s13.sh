set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; tail -1 gpurun_out/bench_n1.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:k_trace_wide -c 1 -o gpurun_out/r2_cw_primary_2048 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extra > /dev/null 2>&1
timeout 600 python -m pytest tests/test_variants_gpu.py tests/test_trace_gpu.py tests/test_harness.py -m gpu -q > gpurun_out/pytest_gpu13.log 2>&1; tail -3 gpurun_out/pytest_gpu13.log
