#!/bin/bash
# developer probe: e2e throughput of the host-buffer entry point under the three transfer strategies
for m in copy zerocopy zerocopy2d; do
  ECB200_HOST_PATH=$m python bench.py --steps 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', d['value'], d['e2e']['value'])"
done
