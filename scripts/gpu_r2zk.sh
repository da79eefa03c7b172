#!/bin/bash
for c in 2 3 4 6 8; do echo "== sweep ctas $c"; AAE_B200_SWEEP_CTAS=$c SUSTAIN_S=0.2 timeout 200 python scripts/step_trace.py 2>&1 | grep "^cold" | cut -c1-330; done
echo "== pubmed"; for c in 2 4; do WL=pubmed AAE_B200_SWEEP_CTAS=$c SUSTAIN_S=0.2 timeout 200 python scripts/step_trace.py 2>&1 | grep "^cold" | cut -c1-330; done
