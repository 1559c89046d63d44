timeout 300 python tools/scratch/standalone.py || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/standalone_launches.csv python tools/scratch/standalone.py > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
