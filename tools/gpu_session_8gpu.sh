# 8-GPU session: the default bench line at N=8 (strong scaling of the one 67.1 M + 67.1 M ray set), BASELINE config 4 (Bistro, CWBVH,
# 536.8 M camera + shadow + diffuse-bounce rays sharded over 8 GPUs, one NCCL broadcast), the one-process group API over 8 real devices
set -x
mkdir -p gpurun_out
nvidia-smi topo -m | head -14 | cut -c1-160 > gpurun_out/topo8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.log; tail -2 gpurun_out/bench_n8.log | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --res 5792 --steps 5 --warmup 3 --no-parity --no-cpu-baseline > gpurun_out/bench_config4_n8.json 2> gpurun_out/bench_config4_n8.log; tail -2 gpurun_out/bench_config4_n8.log | cut -c1-300
timeout 600 python -m pytest tests/test_group_gpu.py -m gpu -q > gpurun_out/pytest_group8.log 2>&1; tail -3 gpurun_out/pytest_group8.log
timeout 600 python tools/group_probe.py > gpurun_out/group8.txt 2>&1; cat gpurun_out/group8.txt
