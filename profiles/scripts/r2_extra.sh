timeout 300 python -m pytest tests/test_dropin_gpu.py -q -s -k "addboth or qsample" 2>&1 | tail -25
