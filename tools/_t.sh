timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_fuzz.py -m gpu -q 2>&1 | tail -2
timeout 100 python tools/time_config.py c2 2>&1 | tail -1
timeout 100 python tools/time_config.py c2 2>&1 | tail -1
timeout 100 python tools/time_config.py c5 2>&1 | tail -1
