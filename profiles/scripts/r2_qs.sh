timeout 300 python -m pytest tests/test_select_gpu.py -q -k "golden" 2>&1 | tail -15
