#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

  launches <launches.csv> <out.md>          per-kernel share of an `ncu --metrics gpu__time_duration.sum` launch list
  kernel   <report.ncu-rep> <out.md>        key metrics of every launch in an `ncu --set full` report
  hot      <report.ncu-rep> <out.md>        stall reasons + hottest source lines (report taken with --import-source on)
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum"]


def launches(path, out):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[start]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        tot[name] += v * scale
        cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list: {path}\n\nserialised, cold-cache per-launch times: compare SHARES, not absolutes\n\n")
        f.write(f"total {T / 1e3:.2f} ms over {sum(cnt.values())} launches\n\n| share | total ms | launches | avg us | kernel |\n|---|---|---|---|---|\n")
        for n, v in tot.most_common(25):
            f.write(f"| {v / T * 100:.2f}% | {v / 1e3:.3f} | {cnt[n]} | {v / cnt[n]:.2f} | `{n}` |\n")


def kernel(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full: {path}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[h.index('Kernel Name')][:90]}  (launch id {r[h.index('ID')]})\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in h:
                    f.write(f"| {k} | {r[h.index(k)]} | {units[h.index(k)]} |\n")
            f.write("\n")


def _f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def hot(path, out, top=30):
    """Stall-reason breakdown + hottest CUDA source lines of the (first) kernel in an `ncu --set full
    --import-source on` report (compile with -lineinfo).  Region rows aggregate 20-line windows of the .cu file."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, v = rows[0], rows[2]
    stalls = [(n.replace("smsp__pcsamp_warps_issue_stalled_", ""), _f(v[i])) for i, n in enumerate(h)
              if n.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in n]
    tot_st = sum(x for _, x in stalls) or 1.0
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    cur, hdr, lines = None, None, []
    for r in csv.reader(src.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            lines.append((cur, int(r[0]), r[1].strip(), _f(r[hdr.index("# Samples")]), _f(r[hdr.index("Instructions Executed")])))
    ts = sum(l[3] for l in lines) or 1.0
    ti = sum(l[4] for l in lines) or 1.0
    region = collections.defaultdict(lambda: [0.0, 0.0])
    for f, ln, _, s_, i_ in lines:
        key = f"{f}:{ln // 20 * 20}-{ln // 20 * 20 + 19}"
        region[key][0] += s_
        region[key][1] += i_
    with open(out, "w") as fo:
        fo.write(f"# ncu source hot spots: {path}\n\n")
        fo.write(f"kernel: `{v[h.index('Kernel Name')][:100]}`, {v[h.index('gpu__time_duration.sum')]} {rows[1][h.index('gpu__time_duration.sum')]}, "
                 f"{int(ti)} warp-instructions, {int(ts)} samples\n\n## stall reasons\n\n| reason | share |\n|---|---|\n")
        for n, x in sorted(stalls, key=lambda t: -t[1])[:10]:
            fo.write(f"| {n} | {100 * x / tot_st:.1f}% |\n")
        fo.write("\n## 20-line regions\n\n| region | samples | instructions |\n|---|---|---|\n")
        for k, (s_, i_) in sorted(region.items(), key=lambda kv: -kv[1][0])[:15]:
            fo.write(f"| {k} | {100 * s_ / ts:.1f}% | {100 * i_ / ti:.1f}% |\n")
        fo.write("\n## lines\n\n| file:line | samples | instructions | source |\n|---|---|---|---|\n")
        for f, ln, text, s_, i_ in sorted(lines, key=lambda l: -l[3])[:top]:
            fo.write(f"| {f}:{ln} | {100 * s_ / ts:.1f}% | {100 * i_ / ti:.1f}% | `{text[:90].replace('|', '/')}` |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "hot": hot}[sys.argv[1]](sys.argv[2], sys.argv[3])
