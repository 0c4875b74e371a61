#!/bin/bash
timeout 1200 compute-sanitizer --tool memcheck python scripts/sanitize_r2d.py > gpurun_out/r2d_sanitize_memcheck.log 2>&1; tail -4 gpurun_out/r2d_sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck python scripts/sanitize_r2d.py > gpurun_out/r2d_sanitize_racecheck.log 2>&1; tail -4 gpurun_out/r2d_sanitize_racecheck.log
