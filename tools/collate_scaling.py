"""Collate bench.py lines measured at N = 1, 2, 4, 8 (gpurun_out/r02_scale_N*.json; N = 1 from the full single-GPU run) into
profiles/r02_scaling.json.  Efficiency = value(N) / (N * value(1)): the north-star split is STRONG scaling (64 restarts in
total)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_json(path):
    with open(path) as f:
        lines = [l for l in f.read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def main():
    src = {1: sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_bench_e.json")}
    for n in (2, 4, 8):
        src[n] = os.path.join(ROOT, "gpurun_out", f"r02_scale_N{n}.json")
    rows, base = [], None
    for n in (1, 2, 4, 8):
        d = last_json(src[n])
        assert d["n_gpus"] == n, (n, d["n_gpus"])
        base = d["value"] if n == 1 else base
        e2e = d["e2e"]
        bo = d.get("bo_iter_sharded") or {}
        rows.append({
            "n_gpus": n, "value_evals_per_s": d["value"], "ms_per_step": d["ms_per_step"],
            "e2e_evals_per_s": e2e["value"], "e2e_ms_per_step": e2e.get("ms_per_step"),
            "e2e_via_forward_autograd_ms": (e2e.get("via_forward_autograd") or {}).get("ms_per_step"),
            "dominant_kernel_frac": d["roofline"]["frac"], "whole_step_frac": d["roofline"].get("whole_step_frac"),
            "weak_evals_per_s": (d.get("weak") or {}).get("value"),
            "bo_iter_sharded_sec": bo.get("sec"), "bo_iter_sharded_phases": bo.get("phases_sec"),
            "clocks": d.get("clocks"), "strong_efficiency_value": d["value"] / (n * base),
        })
    e1 = rows[0]["e2e_evals_per_s"]
    for r in rows:
        r["strong_efficiency_e2e"] = r["e2e_evals_per_s"] / (r["n_gpus"] * e1)
    out = {"what": "bench.py at N = 1, 2, 4, 8 (builder-run, gpurun, end of r02): 64 restarts in total x 1024 MC samples, "
                   "n_train = 1000, strong scaling; e2e = prb_mc_value_grad_host with pinned host buffers", "rows": rows}
    with open(os.path.join(ROOT, "profiles", "r02_scaling.json"), "w") as f:
        json.dump(out, f, indent=1)
    for r in rows:
        print(r["n_gpus"], f'{r["value_evals_per_s"] / 1e6:.2f} M', f'{r["ms_per_step"]:.4f} ms',
              f'eff {r["strong_efficiency_value"]:.3f}', f'e2e {r["e2e_evals_per_s"] / 1e6:.2f} M',
              f'eff {r["strong_efficiency_e2e"]:.3f}', "bo_iter", r["bo_iter_sharded_sec"], "weak", r["weak_evals_per_s"])


if __name__ == "__main__":
    main()
