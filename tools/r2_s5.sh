set -x
mkdir -p gpurun_out
timeout 300 python tools/pcie_probe2.py 24 > gpurun_out/pcie3.txt 2>&1; cat gpurun_out/pcie3.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; tail -3 gpurun_out/bench_n1.log; cat gpurun_out/bench_n1.json
timeout 900 python bench.py --impl reference --steps 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; tail -3 gpurun_out/bench_ref.log; cat gpurun_out/bench_ref.json
