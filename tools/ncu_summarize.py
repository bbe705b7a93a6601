"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/ncu_summarize.py launches gpurun_out/r01_launches_v2.csv profiles/r01_launches.txt
    python tools/ncu_summarize.py full gpurun_out/r01_trgemm_tma_v2.ncu-rep profiles/r01_trgemm_tma_ncu.txt
"""
import collections
import csv
import subprocess
import sys


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    order = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        order.append((name, v))
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if "prb::" in k or "pack_state" in k)
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, aggregated from {src}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches; libprb200 kernels {mine:.1f} us "
                f"({mine / tot:.3f}); the rest is problem setup (torch: Cholesky, triangular solve, data synthesis)\n")
        f.write(f"{'kernel':72s} {'n':>5s} {'total_us':>12s} {'share':>7s} {'share_of_prb':>12s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            sp = f"{v[1] / mine:.3f}" if ("prb::" in k or "pack_state" in k) else ""
            f.write(f"{k[:72]:72s} {v[0]:5d} {v[1]:12.1f} {v[1] / tot:7.3f} {sp:>12s}\n")


def full(src, dst):
    det = subprocess.run(["ncu", "-i", src, "--page", "details"], capture_output=True, text=True).stdout
    keep = ("trgemm", "Duration", "Elapsed Cycles", "SM Active Cycles", "Compute (SM) Throughput", "DRAM Throughput",
            "L2 Hit Rate", "Executed Ipc", "Registers Per Thread", "Dynamic Shared Memory", "Grid Size", "Waves Per SM",
            "highest-utilized", "Tensor (DP)", "One or More Eligible", "No Eligible", "Warp Cycles Per Issued",
            "Achieved Occupancy", "Theoretical Occupancy", "Memory Throughput", "L1/TEX Hit", "Mem Busy")
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
            "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fp64"]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, summarised from {src}\n\n## raw metrics\n")
        for i, h in enumerate(H):
            if any(h == w or h.endswith(w) for w in want) and "pct_of" not in h and "per_second" not in h:
                f.write(f"{h} [{rows[1][i]}]: " + " | ".join(r[i] for r in rows[2:]) + "\n")
        f.write("\n## details page (selected lines)\n")
        for line in det.splitlines():
            if any(k in line for k in keep):
                f.write(line.rstrip() + "\n")
        # stall reasons from the source page
        srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(srcp.splitlines()))
        ks = [i for i, r in enumerate(rr) if r and r[0] == "Kernel Name"]
        f.write("\n## warp stall sampling (share of samples) and instruction mix per kernel\n")
        for kidx, start in enumerate(ks):
            end = ks[kidx + 1] if kidx + 1 < len(ks) else len(rr)
            Hs = rr[start + 1]
            idx = {h: i for i, h in enumerate(Hs)}
            data = [r for r in rr[start + 2:end] if len(r) == len(Hs)]
            sc = [h for h in Hs if h.startswith("stall_") and "Not Issued" not in h]
            tot = {h: sum(int(r[idx[h]] or 0) for r in data) for h in sc}
            S = sum(tot.values()) or 1
            f.write(rr[start][1][:90] + "\n  stalls: " + ", ".join(
                f"{h[6:]} {v / S:.3f}" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]) + "\n")
            mix = collections.Counter()
            for r in data:
                toks = r[idx["Source"]].split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                mix[op] += int(r[idx["Instructions Executed"]] or 0)
            T = sum(mix.values()) or 1
            f.write("  instruction mix: " + ", ".join(f"{op} {v / T:.3f}" for op, v in mix.most_common(8)) + "\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
