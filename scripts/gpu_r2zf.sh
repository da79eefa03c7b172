#!/bin/bash
VD=$PWD/aae-recommender_b200/build/variants
echo "== new"; python scripts/k3_sustained.py 2>&1 | tail -2
echo "== old (HEAD)"; AAE_B200_LIB=$VD/lib_k3x_old.so python scripts/k3_sustained.py 2>&1 | tail -2
echo "== new again"; python scripts/k3_sustained.py 2>&1 | tail -2
