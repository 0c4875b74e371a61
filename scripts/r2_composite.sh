#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_composite.py -m gpu -q -x 2>&1 | grep -v "^$\|Warning" | tail -40
