set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_v6.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bo-iter > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:trgemm_tma --launch-skip 4 -c 2 -f -o gpurun_out/r01_trgemm_tma_v6 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bo-iter > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
