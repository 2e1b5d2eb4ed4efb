set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -12 > gpurun_out/topo2.txt
timeout 300 python tools/pcie_probe2.py 24 > gpurun_out/pcie4.txt 2>&1; cat gpurun_out/pcie4.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.log; cut -c1-1800 gpurun_out/bench_n2.json
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_harness.py -m gpu -q -x -s > gpurun_out/pytest_gpu6.log 2>&1; tail -40 gpurun_out/pytest_gpu6.log
