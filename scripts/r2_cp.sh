#!/bin/bash
GGP_CHOL_CLUSTER_INV=1 python scripts/cluster_probe.py 2>&1 | tail -1
GGP_CHOL_CLUSTER_NO_INV=1 python scripts/cluster_probe.py 2>&1 | tail -1
GGP_CHOL_CLUSTER=0 python scripts/cluster_probe.py 2>&1 | tail -1
