"""Column-tile width A/B on few-column shapes (the per-GPU share of the 8-GPU split at mc = 128): run under PRB_BN=32/64/128."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.argv = [sys.argv[0]]
from tools.sweep import measure  # noqa: E402

for n, b, mc in ((1000, 8, 128), (2000, 8, 128), (4000, 8, 128), (500, 8, 1024), (2000, 8, 1024), (1000, 8, 1024), (200, 8, 1024)):
    r = measure(n, b, mc, steps=10, warmup=3)
    print(os.environ.get("PRB_BN", "auto"), n, b, mc, round(r["ms_per_step"], 4), round(r["roofline_frac"], 3), flush=True)
