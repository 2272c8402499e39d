#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_decode.py -q -m gpu -k "t2v" 2>&1 | tail -15 | cut -c1-250
