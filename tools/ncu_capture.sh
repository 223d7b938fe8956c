#!/bin/bash
# one `ncu --set full` capture of the grid kernel on a 1024r x 128z x 2t C5a grid (a number printed
# under ncu is never a bench value); the plain run comes first and must exit 0
T=${1:-cap}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --nt 2 --no-cpu > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err || exit 1
cat gpurun_out/${T}_plain.json
ncu --set full --clock-control none --import-source on -k regex:lh_grid --launch-skip 1 -c 1 \
    -o gpurun_out/${T}_grid -f python bench.py --steps 1 --warmup 1 --nt 2 --no-cpu > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
