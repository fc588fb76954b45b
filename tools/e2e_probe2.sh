#!/bin/bash
# developer probe: packed strategy sweep (pool threads x chunk x pack jobs per chunk x pack look-ahead)
run() { python bench.py --steps 5 --no-cpu-baseline $2 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['e2e']['value'])"; }
export ECB200_HOST_PATH=packed
for mem in pinned pageable; do
for th in 10 12 14; do
  for ch in 65536 131072; do
    ECB200_HOST_THREADS=$th ECB200_HOST_CHUNK=$ch run "$mem packed threads=$th chunk=$ch" "--e2e-memory $mem"
  done
done
ECB200_HOST_THREADS=12 ECB200_HOST_CHUNK=65536 ECB200_PACK_PARTS=16 run "$mem packed threads=12 chunk=65536 parts=16" "--e2e-memory $mem"
ECB200_HOST_THREADS=12 ECB200_HOST_CHUNK=65536 ECB200_PACK_AHEAD=5 run "$mem packed threads=12 chunk=65536 ahead=5" "--e2e-memory $mem"
ECB200_HOST_THREADS=12 ECB200_HOST_CHUNK=65536 ECB200_PACK_AHEAD=2 run "$mem packed threads=12 chunk=65536 ahead=2" "--e2e-memory $mem"
ECB200_HOST_THREADS=12 ECB200_HOST_CHUNK=32768 ECB200_PACK_PARTS=4 ECB200_PACK_AHEAD=5 run "$mem packed threads=12 chunk=32768 parts=4 ahead=5" "--e2e-memory $mem"
done
