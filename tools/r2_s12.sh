set -x
mkdir -p gpurun_out
cat /sys/kernel/mm/transparent_hugepage/enabled; grep -i huge /proc/meminfo | head -5; dmesg 2>/dev/null | grep -i -E "iommu|DMAR" | head -5; cat /proc/cmdline | tr ' ' '\n' | grep -i iommu
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
TBVH_HOST_HUGE=0 timeout 300 python tools/pcie_probe2.py 24 --short > gpurun_out/pcie_4k.txt 2>&1; grep -E "^local pinned" gpurun_out/pcie_4k.txt
TBVH_HOST_HUGE=1 timeout 300 python tools/pcie_probe2.py 24 --short > gpurun_out/pcie_huge.txt 2>&1; grep -E "^local pinned" gpurun_out/pcie_huge.txt
for h in 0 1; do TBVH_HOST_HUGE=$h timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$h bench.py --gpus 2 --steps 5 --no-parity --no-extra > gpurun_out/bench_n2_huge$h.json 2> /dev/null; python -c "
import json,sys
t=open('gpurun_out/bench_n2_huge$h.json').read(); d=json.loads(t[t.find('{\"metric'):].splitlines()[0]); print('huge $h', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'packed', round(d['e2e']['packed_hits_value']))"; done
