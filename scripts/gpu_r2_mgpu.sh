#!/bin/bash
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -25
