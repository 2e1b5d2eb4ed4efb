set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
timeout 120 ./harness/minimal_b200
timeout 300 ./oracle/_ref/speedtest_b200 data/scenes/cryteksponza.bin 800 600
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.log; tail -3 gpurun_out/bench_b.log; cat gpurun_out/bench_b.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches_r1.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_bvh2 -s 7 -c 2 -o gpurun_out/prof_trace_bvh2_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_build_small|k_bin|k_scatter|k_sweep|k_fragments" -c 12 -o gpurun_out/prof_build_r1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --res 256 > gpurun_out/ncu_build.log 2>&1; tail -2 gpurun_out/ncu_build.log
ls -la gpurun_out/
