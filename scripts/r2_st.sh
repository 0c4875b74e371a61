#!/bin/bash
GGP_CHOL_CLUSTER_INV=1 python scripts/step_times.py 400000 --profile 2>&1 | tail -2
GGP_CHOL_CLUSTER_NO_INV=1 python scripts/step_times.py 400000 --profile 2>&1 | tail -2
