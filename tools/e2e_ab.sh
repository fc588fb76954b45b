#!/bin/bash
# developer probe: A/B of the transfer strategies on ONE box, interleaved (hosts differ from box to box and run to run)
run() { python bench.py --steps ${STEPS:-10} --no-cpu-baseline $2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['e2e']['value'])"; }
for i in 1 2 3; do
  ECB200_HOST_PATH=packed run "pinned packed(12)"
  ECB200_HOST_PATH=compact ECB200_HOST_THREADS=8 run "pinned compact(8)"
  ECB200_HOST_PATH=packed ECB200_HOST_THREADS=14 run "pinned packed(14)"
  ECB200_HOST_PATH=packed run "pageable packed(12)" "--e2e-memory pageable"
done
