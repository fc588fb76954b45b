#!/bin/bash
# developer probe: e2e throughput of the host-buffer entry point: caller memory x transfer strategy x pipeline chunk x pool threads
run() { python bench.py --steps 5 --no-cpu-baseline $2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['e2e']['value'])"; }
ECB200_HOST_PATH=copy run "pinned copy"
ECB200_HOST_PATH=compact run "pinned compact threads=8 chunk=65536"
ECB200_HOST_PATH=copy run "pageable copy" "--e2e-memory pageable"
ECB200_HOST_PATH=compact run "pageable compact" "--e2e-memory pageable"
for th in ${THREADS:-8 12 16}; do
  for ch in ${CHUNKS:-32768 65536}; do
    ECB200_HOST_PATH=packed ECB200_HOST_THREADS=$th ECB200_HOST_CHUNK=$ch run "pageable packed threads=$th chunk=$ch" "--e2e-memory pageable"
  done
done
ECB200_HOST_PATH=packed ECB200_HOST_THREADS=12 ECB200_HOST_CHUNK=32768 run "pinned packed threads=12 chunk=32768"
