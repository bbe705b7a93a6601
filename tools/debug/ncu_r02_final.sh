set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bo-iter --no-extras"
$CMD > gpurun_out/ncu_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
CMDG="python bench.py --workload rosenbrock --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bo-iter --no-extras"
$CMDG > gpurun_out/ncu_plain_g.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kstar_q|trgemm_tma" --launch-skip 6 -c 3 -f -o gpurun_out/r02_generic $CMDG > gpurun_out/ncu_full_g.log 2>&1; echo "ncu generic rc=$?"
