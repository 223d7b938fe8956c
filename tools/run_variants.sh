#!/bin/bash
# debug: tests + bench of the current library, then bench with library variants in tools/micro/lib_*.so,
# then the per-phase clock64 breakdown (tools/libprof.so)
mkdir -p gpurun_out
T=${1:-v}
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('main', d['ms_per_step'], d['value'], d['e2e']['value'])"
echo -n "main, single-layer grid (zfrac 0.97): "
python bench.py --steps 3 --warmup 2 --no-cpu --zfrac 0.97 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
cp unconfined_b200/libunconfined_b200.so /tmp/keep.so
for f in tools/micro/lib_*.so; do
  cp $f unconfined_b200/libunconfined_b200.so
  echo -n "$f: "
  python bench.py --steps 3 --warmup 2 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"
done
python tools/prof_run.py 2>&1 | tee gpurun_out/${T}_phases.txt
cp /tmp/keep.so unconfined_b200/libunconfined_b200.so
