set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/final_r02n_pytest.log 2>&1; tail -3 gpurun_out/final_r02n_pytest.log
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/final_r02n_smoke.log 2>&1; tail -5 gpurun_out/final_r02n_smoke.log
( time python bench.py --impl reference ) > gpurun_out/bench_r02n_ref.json 2> gpurun_out/bench_r02n_ref.err; tail -c 600 gpurun_out/bench_r02n_ref.json; tail -4 gpurun_out/bench_r02n_ref.err
( time python bench.py ) > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err; tail -4 gpurun_out/bench_r02n.err; head -c 400 gpurun_out/bench_r02n.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02n.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_r02n.log 2>&1; tail -2 gpurun_out/launches_r02n.csv
