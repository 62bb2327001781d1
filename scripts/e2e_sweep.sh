#!/bin/bash
# e2e (host buffers) throughput against the host pipeline's chunk size
for c in 512 1024 2048 4096; do
  FSEB200_HOST_CHUNK_BLOCKS=$c timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c', d['e2e']['value'], d['value'])"
done
