#!/bin/bash
# tile-order sweep (PRB_MIXED_ORDER 0 = heavy-first, 1 = mixed except the last group, 2 = mixed everywhere): bench
# ms_per_step, then the per-CTA phase trace of the default order (needs the PRB_TRACE_CTAS build, see cta_trace.py)
for m in 0 1 2; do
  r=$(PRB_MIXED_ORDER=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])")
  echo "mixed $m: $r"
done
if [ -f bo_pr_b200/_lib/libprb200_trace.so ]; then
  for m in 0 1; do
    echo "trace, mixed $m"
    PRB_MIXED_ORDER=$m PRB_LIB_NAME=libprb200_trace.so python tools/debug/cta_trace.py
  done
fi
