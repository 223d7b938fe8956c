#!/bin/bash
# debug: A/B timing of library variants (tools/micro/lib_*.so), two passes in alternating order
cp unconfined_b200/libunconfined_b200.so /tmp/keep.so
cp /tmp/keep.so tools/micro/lib_0_main.so
for pass in 1 2; do
for f in tools/micro/lib_*.so; do
  cp $f unconfined_b200/libunconfined_b200.so
  echo -n "pass $pass $f: "
  python bench.py --steps 4 --warmup 3 --no-cpu $EXTRA 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['clocks']['sm_mhz'], d['clocks'].get('power_w_max'))"
done
done
cp /tmp/keep.so unconfined_b200/libunconfined_b200.so
rm tools/micro/lib_0_main.so
