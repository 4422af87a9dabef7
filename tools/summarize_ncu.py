"""Turns the ncu outputs of tools/evidence*.sh into the markdown summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/r01b_launches.csv   > profiles/r01b_launches_summary.md
  python tools/summarize_ncu.py full     gpurun_out/r01b_kernels_raw.csv > profiles/r01b_kernels_ncu_summary.md
"""
import collections
import csv
import re
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"^void ", "", row["Kernel Name"])
        name = re.sub(r"\(.*", "", name)
        tot[name][0] += 1
        tot[name][1] += v
        n += 1
    T = sum(v[1] for v in tot.values())
    ours = sum(v[1] for k, v in tot.items() if "b200mm::" in k)
    gemm = sum(v[1] for k, v in tot.items() if "gemm_tcgen05" in k or "splitk" in k)
    print(f"# ncu launch list `{path}`: {n} launches, {T / 1e3:.1f} ms of kernel time (cold-cache, serialised replays: use the SHARES)\n")
    print(f"b200mm kernels: {100 * ours / T:.2f} % of kernel time; tcgen05 GEMM (+ split-K reduce): {100 * gemm / T:.2f} %\n")
    print("| kernel | launches | total ms | share % |\n|---|---|---|---|")
    for k, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if us / T < 2e-4:
            continue
        print(f"| `{k[:110]}` | {c} | {us / 1e3:.2f} | {100 * us / T:.2f} |")


def full(path):
    r = csv.reader(open(path))
    hdr = next(r)
    units = next(r)
    col = {h: i for i, h in enumerate(hdr)}
    to_gb = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}

    def val(row, i):  # metric value normalised to GB / us where the unit row says bytes / time
        v = float(row[i].replace(",", ""))
        return v * to_gb.get(units[i], 1.0) * (to_us.get(units[i], 1.0) if units[i] in to_us else 1.0)

    def find(sub):
        for h, i in col.items():
            if h.endswith(sub):
                return i
        return None

    want = [("time us", "gpu__time_duration.sum"), ("dram rd GB", "dram__bytes_read.sum"), ("dram wr GB", "dram__bytes_write.sum"),
            ("dram TB/s (rd+wr)/t", None), ("tensor pipe %", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"),
            ("issue active %", "sm__inst_issued.avg.pct_of_peak_sustained_active"), ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ("regs", "launch__registers_per_thread"), ("warp inst (M)", "smsp__inst_executed.sum"), ("XU pipe %", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")]
    idx = [(n, find(m) if m else -1) for n, m in want]
    print(f"# ncu --set full --clock-control none, one launch per hot kernel (tools/prof_kernels.py) — `{path}`\n")
    print("| kernel | grid | " + " | ".join(n for n, _ in idx) + " |\n|---|---|" + "---|" * len(idx))
    for row in r:
        name = re.sub(r"\(.*", "", re.sub(r"^void ", "", row[col["Kernel Name"]]))[:70]
        vals = []
        for n, i in idx:
            if i == -1:
                t = val(row, find("gpu__time_duration.sum"))
                b = val(row, find("dram__bytes_read.sum")) + val(row, find("dram__bytes_write.sum"))
                vals.append(f"{b / t * 1e3:.2f}")  # GB per us = 1e3 TB/s
                continue
            if i is None or row[i] == "":
                vals.append("n/a")
                continue
            v = val(row, i)
            vals.append(f"{v / 1e6:.1f}" if n.startswith("warp inst") else f"{v:.3g}" if v < 100 else f"{v:.1f}")
        print(f"| `{name}` | {row[col['Grid Size']]} | " + " | ".join(vals) + " |")


def traffic(path):
    """JSON from the long-format CSV of `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv`:
    first captured launch of each kernel -> measured time and DRAM bytes."""
    import json

    lines = [l for l in open(path) if not l.startswith("==")]
    to_b = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    to_us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    per_id, order = {}, []
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", re.sub(r"^void ", "", row["Kernel Name"])).replace("b200mm::", "")
        key = row["ID"]
        if key not in per_id:
            per_id[key] = {"name": name, "grid": row["Grid Size"]}
            order.append(key)
        v = float(row["Metric Value"].replace(",", ""))
        m, u = row["Metric Name"], row["Metric Unit"]
        if m == "gpu__time_duration.sum":
            per_id[key]["time_us"] = round(v * to_us.get(u, 1.0), 2)
        elif m == "dram__bytes_read.sum":
            per_id[key]["dram_read_bytes"] = int(v * to_b.get(u, 1.0))
        elif m == "dram__bytes_write.sum":
            per_id[key]["dram_write_bytes"] = int(v * to_b.get(u, 1.0))
    out = {}
    for key in order:
        d = per_id[key]
        if d["name"] in out or "dram_read_bytes" not in d:
            continue
        d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0)
        out[d.pop("name")] = d
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 263168
    # [M, N, K] of the GEMM launches tools/prof_kernels.py makes, by kernel instantiation (flavor = epilogue bits, gemm_tcgen05.cu flavor_bits)
    shapes = {"gemm_tcgen05_kernel<0, 0, 0, 2, 67>": [T, 4096, 1024],   # c_fc forward: bias + QuickGELU + pre-activation copy
              "gemm_tcgen05_kernel<0, 1, 0, 2, 72>": [T, 4096, 1024],   # dgrad through the activation (x act'(u))
              "gemm_tcgen05_kernel<0, 1, 0, 2, 74>": [T, 4096, 1024],   # ... also emitting act(u): the training step's dominant launch
              "gemm_tcgen05_kernel<0, 0, 0, 2, 1>": [T, 3072, 1024],    # qkv projection: bias
              "gemm_tcgen05_kernel<0, 1, 0, 2, 0>": [T, 1024, 1024],    # plain dgrad through the out-projection
              "gemm_tcgen05_kernel<0, 0, 0, 2, 5>": [T, 1024, 4096]}    # c_proj forward: bias + residual
    print(json.dumps({"source": path, "how": f"PROF_ROWS={T} ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                      "--clock-control none python tools/prof_kernels.py 1 (one launch per kernel)",
                      "shapes": {k: v for k, v in shapes.items() if k in out}, "kernels": out}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
