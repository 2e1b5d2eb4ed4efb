set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
for v in 0 1 2; do TBVH_TRACE_VARIANT=$v timeout 300 python tools/quick_perf.py sponza 512 2>&1 | grep -E "GPU build|primary|shadow|diffuse"; done
timeout 300 python tools/quick_perf.py sponza 512 cwbvh 2>&1 | grep -E "CWBVH|primary|shadow|diffuse"
for t in 64 128; do TBVH_SMALL_T=$t timeout 300 python tools/quick_perf.py sponza 256 2>&1 | grep -E "GPU build"; done
TBVH_HOST_PATH=copy2d TBVH_H2D_SPLIT=1 timeout 300 python tools/pcie_probe.py
TBVH_HOST_PATH=copy2d TBVH_H2D_SPLIT=2 timeout 300 python tools/pcie_probe.py | head -2
TBVH_HOST_PATH=copy2d TBVH_H2D_SPLIT=4 timeout 300 python tools/pcie_probe.py | head -2
TBVH_HOST_PATH=zerocopy timeout 300 python tools/pcie_probe.py | head -2
timeout 400 python bench.py --steps 5 --warmup 3 --layout cwbvh --no-cpu-baseline > gpurun_out/bench_cw.json 2> gpurun_out/bench_cw.log; tail -3 gpurun_out/bench_cw.log; cat gpurun_out/bench_cw.json
