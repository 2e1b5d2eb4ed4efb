# One gpurun call that refreshes every measured artefact under profiles/ (about 6-7 GPU-minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session_full.sh'
# (remove the two Bistro lines from .gpurunignore first if the Bistro tests / timings are wanted: +2 x 68 MB to push)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log
timeout 600 python bench.py --tree sah --no-cpu-baseline > gpurun_out/bench_n1_sah.json 2> /dev/null
# launch list of the bench command + one full capture of the dominant kernel (traffic for profiles/traffic.json)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_bvh2 -s 3 -c 1 -o gpurun_out/trace python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
# builders, SBVH phases, TLAS, SAH costs, the C++ harness
timeout 300 python tools/quick_build.py sponza bunny > gpurun_out/build.log 2>&1; tail -2 gpurun_out/build.log
timeout 300 python tools/quick_hq.py bunny sponza > gpurun_out/hq.log 2>&1; tail -2 gpurun_out/hq.log
TBVH_HQ_PROFILE=2 timeout 120 python tools/hq_once.py sponza > gpurun_out/hq_profile.log 2>&1
timeout 300 python tools/quick_tlas.py 32 1024 > gpurun_out/tlas.log 2>&1; tail -5 gpurun_out/tlas.log
timeout 300 python tools/sah_check.py > gpurun_out/sah.log 2>&1; tail -6 gpurun_out/sah.log
(cd oracle/_ref && timeout 400 ./speedtest_b200 ../../data/scenes/cryteksponza.bin) > gpurun_out/speedtest_b200.log 2>&1; echo "harness rc=$?"; tail -8 gpurun_out/speedtest_b200.log
