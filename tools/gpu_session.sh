set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_build_small -c 1 -o gpurun_out/prof_build_small python -c "
import sys; sys.path.insert(0,'.')
from tinybvh_b200 import api, scenes
v,_ = scenes.load_scene('bistro'); e = api.BVH().Build(v); print(e.info().build_ms)
" > gpurun_out/ncu_bs.log 2>&1; tail -2 gpurun_out/ncu_bs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bistro_build3.csv python -c "
import sys; sys.path.insert(0,'.')
from tinybvh_b200 import api, scenes
v,_ = scenes.load_scene('bistro'); e = api.BVH().Build(v); print(e.info().build_ms)
" > gpurun_out/ncu_bistro_build.log 2>&1; tail -2 gpurun_out/ncu_bistro_build.log
