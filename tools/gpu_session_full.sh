# One gpurun call that refreshes every measured artefact under profiles/ (about 10 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_session_full.sh'
# then here:  python tools/collect_profiles.py r2
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
timeout 900 python bench.py --impl reference --steps 10 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log
timeout 900 python bench.py --layout bvh --no-cpu-baseline --steps 10 > gpurun_out/bench_n1_bvh.json 2> /dev/null
timeout 900 python bench.py --scene sponza --layout bvh --res 1024 --no-cpu-baseline --steps 10 > gpurun_out/bench_sponza_bvh.json 2> /dev/null
timeout 900 python bench.py --scene sponza --layout cwbvh --res 1024 --no-cpu-baseline --steps 10 > gpurun_out/bench_sponza_cwbvh.json 2> /dev/null
# BASELINE config 5: lucy + dragon x 29 (10.1 M triangles), binned-SAH build on the GPU, 268 M camera + 268 M shadow + 268 M bounce rays
timeout 1200 python bench.py --scene lucy_dragon_x29 --layout bvh --tree sah --res 4096 --steps 3 --no-cpu-baseline --parity-rays 262144 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.log
# launch list of the bench command + one full capture of the dominant kernel (traffic for profiles/traffic.json)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --res 1024 --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-extra > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 2 -c 1 -o gpurun_out/r2_cw_primary python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets primary > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 3 -c 1 -o gpurun_out/r2_cw_shadow python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets primary,shadow > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_trace_wide -c 1 -o gpurun_out/r2_cw_primary_2048 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-extra > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_large_phase -c 1 -o gpurun_out/r2_large_phase python tools/quick_build.py sponza > /dev/null 2>&1
timeout 300 python tools/quick_build.py sponza bunny bistro lucy_dragon_x29 > gpurun_out/build.log 2>&1; tail -4 gpurun_out/build.log
timeout 300 python tools/quick_hq.py bunny sponza bistro > gpurun_out/hq.log 2>&1; tail -3 gpurun_out/hq.log
ls -la gpurun_out | tail -20
