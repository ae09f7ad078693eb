timeout 600 python -m pytest tests/test_slab_gpu.py tests/test_visual.py -m gpu -x -q -k "pressure_range_and_frame or real_time" 2>&1 | tail -15
