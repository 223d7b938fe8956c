#!/bin/bash
# One GPU-box session: tests, error budget, bench, launch list and ncu captures.  Everything a
# later step needs is written under gpurun_out/ (merged back by gpurun).  usage: gpu_session.sh TAG [steps...]
T=${1:-s}; shift
STEPS=${@:-"test budget bench c5b ncu_point ncu_grid"}
mkdir -p gpurun_out
for S in $STEPS; do
case $S in
test)   ( time timeout 1700 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1; tail -15 gpurun_out/${T}_pytest.log ;;
testall) ( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1; tail -40 gpurun_out/${T}_pytest.log ;;
budget) timeout 900 python tools/error_budget.py run gpurun_out/${T}_error_budget.txt > gpurun_out/${T}_budget.log 2>&1; tail -40 gpurun_out/${T}_budget.log ;;
bench)  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err ;;
benchq) timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err ;;
ref)    timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cat gpurun_out/${T}_bench_ref.json ;;
c5b)    timeout 600 python tools/bench_c5b.py 17 3 > gpurun_out/${T}_c5b.json 2> gpurun_out/${T}_c5b.err; cat gpurun_out/${T}_c5b.json; tail -3 gpurun_out/${T}_c5b.err ;;
variants) timeout 1500 python tools/variants.py run gpurun_out/${T}_variants.txt > gpurun_out/${T}_variants.log 2>&1; cat gpurun_out/${T}_variants.txt ;;
ngpu)   timeout 600 python tools/bench_ngpu.py gpurun_out/${T}_strong_ngpu.json > gpurun_out/${T}_ngpu.log 2>&1; tail -12 gpurun_out/${T}_ngpu.log ;;
parity) timeout 1500 python tools/parity_report.py --out gpurun_out/${T}_parity.txt > gpurun_out/${T}_parity.log 2>&1; tail -70 gpurun_out/${T}_parity.txt ;;
decks)  timeout 600 python tools/bench_decks.py > gpurun_out/${T}_decks.txt 2>&1; cat gpurun_out/${T}_decks.txt ;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1; tail -2 gpurun_out/${T}_ncu_launch.log ;;
ncu_point) timeout 900 ncu --set full --clock-control none --import-source on -k regex:lh_point --launch-skip 1 -c 1 -o gpurun_out/${T}_point -f python tools/bench_c5b.py 15 1 > gpurun_out/${T}_ncu_point.log 2>&1; tail -2 gpurun_out/${T}_ncu_point.log ;;
ncu_grid) python -c "import bench; print(bench.sass_sha256('unconfined_b200/libunconfined_b200.so'))" > gpurun_out/${T}_sass.sha; timeout 900 ncu --set full --clock-control none --import-source on -k regex:lh_grid8 --launch-skip 1 -c 1 -o gpurun_out/${T}_grid -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_grid.log 2>&1; tail -2 gpurun_out/${T}_ncu_grid.log ;;
esac
done
