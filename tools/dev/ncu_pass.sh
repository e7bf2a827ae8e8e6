set -e
python bench.py --precision f16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_f16.csv python bench.py --precision f16 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
