#!/bin/bash
rm -f gpurun_out/headline_parity.json
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$\|Warning" | tail -80
cat gpurun_out/headline_parity.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    print(k, 'jitter', v['jitter'])
    for kk,e in v['errors'].items(): print('   ', kk, {a: float('%.2e'%b) for a,b in e.items()})
"
