#!/bin/bash
VD=$PWD/aae-recommender_b200/build/variants
echo "== new"; timeout 200 python scripts/step_trace.py 2>&1 | tail -5 | cut -c1-400
echo "== old"; AAE_B200_LIB=$VD/lib_k3x_old.so timeout 200 python scripts/step_trace.py 2>&1 | tail -5 | cut -c1-400
