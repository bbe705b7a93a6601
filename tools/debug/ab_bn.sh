for BN in 0 33; do
  for R in 64 8; do
    echo "== PRB_BN=$BN restarts=$R"
    PRB_BN=$BN python bench.py --restarts $R --steps 20 --warmup 5 --no-cpu-baseline --no-bo-iter --no-extras --no-e2e 2>&1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], {k:round(v['ms_per_step'],4) for k,v in d['kernel_breakdown'].items()})"
  done
done
PRB_BN=33 python -m pytest tests/test_gpu_synthetic.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
