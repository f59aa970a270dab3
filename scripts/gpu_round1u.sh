#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backend_plugin.py tests/test_gpu_gpt2.py -q -m gpu --timeout 600 2>&1 | tail -5
