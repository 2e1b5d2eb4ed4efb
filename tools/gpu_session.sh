set -x
: > .gpurunignore.tmp
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
timeout 300 python tools/quick_perf.py sponza 1024 cwbvh 2>&1 | grep -E "CWBVH|primary|shadow|diffuse"
timeout 300 python tools/quick_perf.py bistro 1024 cwbvh 2>&1 | grep -E "CWBVH|primary|shadow|diffuse"
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.log; tail -3 gpurun_out/bench_d.log; cat gpurun_out/bench_d.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2>/dev/null; cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -1 gpurun_out/ncu_launch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_bvh2 -s 7 -c 2 -o gpurun_out/prof_trace_bvh2_r1b python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
