#!/bin/bash
# developer probe: e2e throughput of the host-buffer entry point: transfer strategy x pipeline chunk x scatter threads
run() { python bench.py --steps 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['e2e']['value'])"; }
ECB200_HOST_PATH=copy run "copy"
for th in 2 4 8 12; do
  for ch in 32768 65536 131072; do
    ECB200_HOST_THREADS=$th ECB200_HOST_CHUNK=$ch run "compact threads=$th chunk=$ch"
  done
done
