#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_rnn_parity.py -m gpu -q --timeout=800 -k "large_batch" 2>&1 | tail -3
