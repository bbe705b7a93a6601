set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bo-iter --no-extras"
$CMD > gpurun_out/ncu_plain64.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:trgemm_tma --launch-skip 4 -c 2 -f -o gpurun_out/r02_trgemm_b64 $CMD > gpurun_out/ncu_full64.log 2>&1; echo "ncu full64 rc=$?"
$CMD --restarts 8 > gpurun_out/ncu_plain8.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:trgemm_tma --launch-skip 4 -c 2 -f -o gpurun_out/r02_trgemm_b8 $CMD --restarts 8 > gpurun_out/ncu_full8.log 2>&1; echo "ncu full8 rc=$?"
$CMD > /dev/null 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
