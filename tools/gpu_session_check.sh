set -x
mkdir -p gpurun_out
# new / changed tests first (their verdict survives a clamped call), then the whole GPU suite, smoke(), and the two-level probe
timeout 300 python -m pytest tests/test_tlas_gpu.py -m gpu -q > gpurun_out/pytest_gpu_tlas.log 2>&1; tail -3 gpurun_out/pytest_gpu_tlas.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python tools/quick_tlas.py > gpurun_out/quick_tlas.txt 2>&1; tail -4 gpurun_out/quick_tlas.txt
if [ -n "$WITH_BENCH" ]; then timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; tail -1 gpurun_out/bench_n1.log | cut -c1-100; fi
