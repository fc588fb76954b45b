#!/bin/bash
# developer probe: e2e throughput of the host-buffer entry point vs pipeline chunk size
for ch in 16384 32768 65536 131072 262144; do
  ECB200_HOST_CHUNK=$ch python bench.py --steps 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $ch', d['value'], d['e2e']['value'])"
done
