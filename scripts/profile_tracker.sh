NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-ba --no-cpu-baseline > gpurun_out/launch_bench.log 2>&1
$NCU --set full --import-source on --profile-from-start off -f -o gpurun_out/tracker \
    python bench.py --steps 2 --warmup 3 --no-ba --no-cpu-baseline --ncu-step > gpurun_out/ncu_tracker.log 2>&1
ls -la gpurun_out/tracker.ncu-rep; wc -l gpurun_out/launches.csv
