timeout 175 python -m pytest tests/test_select_gpu.py tests/test_dropin_gpu.py -q -x 2>&1 | tail -6
