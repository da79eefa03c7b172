#!/bin/bash
# one call: the N=1 record (tests, smoke, both bench arms), the profiling pass, the K3 tunables sweep
bash scripts/gpu_record.sh 2>&1 | tail -25
bash scripts/gpu_profile.sh 2>&1 | tail -30
bash scripts/k3_tunables.sh pf2 pf6 pf8 nomv sl0 sl200 g1l nopin 2>&1 | tail -12
