set -x
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 5 --warmup 3 --no-parity --no-extra > gpurun_out/bench_n4_$tag.json 2> gpurun_out/bench_n4_$tag.log; python -c "
import json
t=open('gpurun_out/bench_n4_$tag.json').read(); d=json.loads(t[t.find('{\"metric'):].splitlines()[0]); print('$tag', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e']['packed_hits_value']), 'identical', d['e2e']['results_identical_to_device_path'])"; }
run base X=1
run fullline TBVH_D2H_MODE=1
run spread TBVH_BENCH_NUMA_SPREAD=1
run spread_fullline TBVH_BENCH_NUMA_SPREAD=1 TBVH_D2H_MODE=1
timeout 300 python -m pytest tests/test_variants_gpu.py -m gpu -q -k host_path > gpurun_out/pytest15.log 2>&1; tail -2 gpurun_out/pytest15.log
