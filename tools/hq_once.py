import sys; sys.path.insert(0,'/root/repo')
from tinybvh_b200 import api, scenes
v,_ = scenes.load_scene(sys.argv[1] if len(sys.argv)>1 else 'sponza')
e = api.BVH().BuildHQ(v); print(e.info().build_ms)
