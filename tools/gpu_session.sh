set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for m in 0 1 2 3; do TBVH_D2H_MODE=$m timeout 300 python tools/pcie_probe.py 2>&1 | tail -2; done
TBVH_HOST_PATH=zerocopy TBVH_D2H_MODE=2 timeout 300 python tools/pcie_probe.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_bistro_gpu.py -m gpu -q -s 2>&1 | tail -8
timeout 600 python tools/quick_perf.py bistro 1024 2>&1 | grep -E "GPU build|primary|shadow|diffuse|host path"
timeout 600 python tools/quick_perf.py bistro 1024 cwbvh 2>&1 | grep -E "CWBVH|primary|shadow|diffuse"
timeout 900 python bench.py --scene bistro --layout cwbvh --steps 3 --warmup 3 > gpurun_out/bench_bistro_cw.json 2> gpurun_out/bench_bistro_cw.log; tail -3 gpurun_out/bench_bistro_cw.log; cat gpurun_out/bench_bistro_cw.json
timeout 900 python bench.py --scene bistro --layout bvh --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bistro_bvh.json 2> gpurun_out/bench_bistro_bvh.log; tail -3 gpurun_out/bench_bistro_bvh.log; cat gpurun_out/bench_bistro_bvh.json
