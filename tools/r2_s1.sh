# round-2 session 1: diagnostics (Bistro parity at size, CWBVH vs BVH2 on Bistro with counters, ncu full on the CWBVH kernel, host topology)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; numactl -H >> gpurun_out/topo.txt 2>&1; lscpu | head -30 >> gpurun_out/topo.txt; free -g >> gpurun_out/topo.txt
./tools/ubench/subwarp > gpurun_out/subwarp.txt 2>&1; cat gpurun_out/subwarp.txt
timeout 900 python -m pytest tests/test_bistro_gpu.py -m gpu -q -x -s > gpurun_out/pytest_bistro.log 2>&1; tail -8 gpurun_out/pytest_bistro.log
timeout 300 python tools/trace_once.py bistro 1024 cwbvh --stats > gpurun_out/t_bistro_cwbvh.txt 2>&1; cat gpurun_out/t_bistro_cwbvh.txt
timeout 300 python tools/trace_once.py bistro 1024 bvh --stats > gpurun_out/t_bistro_bvh.txt 2>&1; cat gpurun_out/t_bistro_bvh.txt
timeout 300 python tools/trace_once.py sponza 1024 cwbvh --stats > gpurun_out/t_sponza_cwbvh.txt 2>&1; cat gpurun_out/t_sponza_cwbvh.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_cwbvh -s 2 -c 1 -o gpurun_out/cw_primary python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets primary > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_cwbvh -s 3 -c 1 -o gpurun_out/cw_diffuse python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets diffuse > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_bvh2 -s 2 -c 1 -o gpurun_out/b2_primary python tools/trace_once.py bistro 1024 bvh --reps 1 --sets primary > /dev/null 2>&1
ls -la gpurun_out
