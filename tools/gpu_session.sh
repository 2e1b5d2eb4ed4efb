set -x
timeout 600 python tools/quick_build.py sponza bistro lucy_dragon_x29 2>&1 | grep -E "build"
timeout 900 python bench.py --scene lucy_dragon_x29 --layout bvh --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ld.json 2> gpurun_out/bench_ld.log; tail -3 gpurun_out/bench_ld.log; cut -c1-1500 gpurun_out/bench_ld.json
