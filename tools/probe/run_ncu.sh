python tools/profile_target.py --workload 16384x4096 --steps 4 --warmup 2 > gpurun_out/r2_ncu_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lbm_stream -s 2 -c 1 -o gpurun_out/r2_ncu_stream python tools/profile_target.py --workload 16384x4096 --steps 4 --warmup 2 > gpurun_out/r2_ncu_log.txt 2>&1
cat gpurun_out/r2_ncu_plain.log
python bench.py --steps 20 --warmup 4 --no-extra --repeats 2 > gpurun_out/r2_launch_plain.json 2> gpurun_out/r2_launch_plain.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_16384.csv python bench.py --steps 20 --warmup 4 --no-extra --repeats 2 > gpurun_out/r2_launch_ncu.json 2> gpurun_out/r2_launch_ncu.err
wc -l gpurun_out/r2_launches_bench_16384.csv
