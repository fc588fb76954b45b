#!/bin/bash
# developer probe: A/B of host-path settings on ONE box, interleaved (hosts differ from box to box and run to run)
run() { python bench.py --steps ${STEPS:-10} --no-cpu-baseline $2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['e2e']['value'])"; }
for i in 1 2 3; do
  ECB200_HOST_PATH=packed run "pinned packed nt"
  ECB200_HOST_PATH=packed ECB200_PACK_NT=0 run "pinned packed plain-stores"
  ECB200_HOST_PATH=packed ECB200_HOST_CHUNK=32768 ECB200_PACK_NT=0 ECB200_PACK_AHEAD=4 run "pinned packed plain-stores chunk=32768 ahead=4"
  ECB200_HOST_PATH=packed ECB200_HOST_CHUNK=32768 ECB200_PACK_AHEAD=4 run "pinned packed nt chunk=32768 ahead=4"
done
