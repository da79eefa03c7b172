#!/bin/bash
for v in 0 1 2 3; do timeout 45 python -u scripts/probe_mn_major.py $v 2>&1 | tail -5; echo "variant $v rc=$?"; done
