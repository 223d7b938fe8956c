#!/bin/bash
# debug: time bench.py with library variants that skip phases of lh_grid4_kernel
cp unconfined_b200/libunconfined_b200.so /tmp/keep.so
for f in tools/micro/lib_*.so; do
  cp $f unconfined_b200/libunconfined_b200.so
  echo -n "$f: "
  python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done
cp /tmp/keep.so unconfined_b200/libunconfined_b200.so
