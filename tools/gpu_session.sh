set -x
timeout 900 python -m pytest tests/test_build_gpu.py tests/test_convert_gpu.py tests/test_bistro_gpu.py tests/test_variants_gpu.py -m gpu -q 2>&1 | tail -4
timeout 600 python tools/quick_build.py sponza bistro lucy_dragon_x29 2>&1 | grep -E "build"
