timeout 300 python -m pytest tests/test_dropin_gpu.py -q -k "vectorised_comparison and (DET or QSAMPLE or ADDBOTH)" 2>&1 | tail -25
