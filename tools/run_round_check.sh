mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/s2_pytest.log 2>&1
python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s2_bench_ref.json 2> gpurun_out/s2_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/s2_ncu_launch.log 2>&1
cp unconfined_b200/libunconfined_b200.so /tmp/keep.so
python tools/prof_run.py > gpurun_out/s2_phases.txt 2>&1
cp /tmp/keep.so unconfined_b200/libunconfined_b200.so
tail -3 gpurun_out/s2_pytest.log; cat gpurun_out/s2_bench.json; cat gpurun_out/s2_bench_ref.json; cat gpurun_out/s2_phases.txt
