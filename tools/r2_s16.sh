set -x
mkdir -p gpurun_out
timeout 300 python tools/pcie_probe2.py 24 --short > gpurun_out/pcie_d2h1.txt 2>&1; grep -E "^local" gpurun_out/pcie_d2h1.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; tail -1 gpurun_out/bench_n1.log | cut -c1-200
timeout 600 python -m pytest tests/test_variants_gpu.py tests/test_tlas_gpu.py tests/test_group_gpu.py -m gpu -q > gpurun_out/pytest16.log 2>&1; tail -2 gpurun_out/pytest16.log
