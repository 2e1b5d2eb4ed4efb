# round-2 session 2: new CWBVH traversal (parity + speed), full GPU test-suite, pipe micro-probe, PCIe / NUMA probe
set -x
mkdir -p gpurun_out
./tools/ubench/pipes > gpurun_out/pipes.txt 2>&1; cat gpurun_out/pipes.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/trace_once.py bistro 1024 cwbvh --stats > gpurun_out/t_bistro_cwbvh.txt 2>&1; cat gpurun_out/t_bistro_cwbvh.txt
timeout 300 python tools/trace_once.py sponza 1024 cwbvh --stats > gpurun_out/t_sponza_cwbvh.txt 2>&1; cat gpurun_out/t_sponza_cwbvh.txt
timeout 300 python tools/pcie_probe2.py 24 > gpurun_out/pcie2.txt 2>&1; cat gpurun_out/pcie2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_wide -s 2 -c 1 -o gpurun_out/cw2_primary python tools/trace_once.py bistro 1024 cwbvh --reps 1 --sets primary > /dev/null 2>&1
ls -la gpurun_out
