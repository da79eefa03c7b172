#!/bin/bash
for G in 8 16 32; do
echo "== W1 groups $G"
AAE_B200_W1_GROUPS=$G python bench.py --workload mpd --no-extra --no-cpu --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('MPD value %.0f ms %.4f sustained %.4f e2e %.0f'%(d['value'],d['ms_per_step'],d['sustained']['ms_per_step'],d['e2e']['value']))"
AAE_B200_W1_GROUPS=$G python bench.py --workload pubmed --no-extra --no-cpu --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('PubMed value %.0f ms %.4f sustained %.4f e2e %.0f'%(d['value'],d['ms_per_step'],d['sustained']['ms_per_step'],d['e2e']['value']))"
done
