set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_perf.py sponza 1024 2>&1 | grep -E "GPU build|primary|shadow|diffuse"
timeout 600 python tools/quick_perf.py bistro 1024 2>&1 | grep -E "GPU build|primary|shadow|diffuse"
TBVH_HOST_PATH=zerocopy TBVH_D2H_MODE=0 timeout 300 python tools/pcie_probe.py 2>&1 | tail -2
TBVH_HOST_PATH=zerocopy TBVH_D2H_MODE=1 timeout 300 python tools/pcie_probe.py 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.log; tail -3 gpurun_out/bench_c.log; cat gpurun_out/bench_c.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bistro_build2.csv python -c "
import sys; sys.path.insert(0,'.')
from tinybvh_b200 import api, scenes
v,_ = scenes.load_scene('bistro'); e = api.BVH().Build(v); print(e.info().build_ms)
" > gpurun_out/ncu_bistro_build.log 2>&1; tail -2 gpurun_out/ncu_bistro_build.log
