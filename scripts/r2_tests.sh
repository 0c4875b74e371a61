#!/bin/bash
rm -f gpurun_out/headline_parity.json
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$\|Warning" | tail -40
