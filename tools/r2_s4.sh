set -x
mkdir -p gpurun_out
timeout 300 python tools/pcie_probe2.py 24 > gpurun_out/pcie2.txt 2>&1; cat gpurun_out/pcie2.txt
timeout 300 python tools/trace_once.py bistro 1024 cwbvh > gpurun_out/t_bistro_cwbvh_oct.txt 2>&1; cat gpurun_out/t_bistro_cwbvh_oct.txt
TBVH_TRACE_VARIANT=0 timeout 300 python tools/trace_once.py bistro 1024 cwbvh > gpurun_out/t_bistro_cwbvh_gen.txt 2>&1; cat gpurun_out/t_bistro_cwbvh_gen.txt
