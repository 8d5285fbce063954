timeout 100 python -m pytest tests/test_dropin_gpu.py -q -s -k "comparison_runs and det" 2>&1 | tail -4
