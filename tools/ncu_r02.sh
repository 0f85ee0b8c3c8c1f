# launch list of the routed kernels (config 5 at 1e9 points)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_route -c 12 --csv --log-file gpurun_out/launches_routed.csv python tools/bench_configs.py --configs 5 > gpurun_out/ncu_routed.log 2>&1
# full capture of the bin kernel
ncu --set full --clock-control none --import-source on -k regex:k_route_bin -s 2 -c 1 -o gpurun_out/prof_route_bin python tools/bench_configs.py --configs 5 > gpurun_out/ncu_routed2.log 2>&1
# traffic of the headline kernels at n = 1e9 (bench, 1 step)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_points_priv_tight -s 4 -c 4 --csv --log-file gpurun_out/traffic_k2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs --no-strong --no-parity > gpurun_out/ncu_k2.log 2>&1
tail -3 gpurun_out/launches_routed.csv; tail -3 gpurun_out/traffic_k2.csv
