#!/bin/bash
GGP_I8_TIMELINE=2 python scripts/prof_one_eval_i8.py 131072 2>&1 | grep -B1 -A22 "epilogue [01]," | cut -c1-200 | head -120
