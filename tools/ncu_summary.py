"""Compact summaries of ncu exports (run where ncu is installed; no GPU needed).
   python tools/ncu_summary.py launches gpurun_out/launches.csv
   python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep
   python tools/ncu_summary.py source gpurun_out/prof.ncu-rep [window]"""
import csv, io, subprocess, sys, collections

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "dram__bytes_read.sum.per_second", "smsp__warps_eligible.avg.per_cycle_active"]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0][-70:]
        v = float(r[iv].replace(",", ""))
        unit = r[hdr.index("Metric Unit")]
        v = v / 1e6 if unit == "ns" else (v / 1e3 if unit == "us" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':72s} {'n':>4s} {'ms':>10s} {'share':>7s}")
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:72s} {n:4d} {ms:10.3f} {100 * ms / tot:6.1f}%")
    print(f"{'total':72s} {sum(a[0] for a in agg.values()):4d} {tot:10.3f}")


def raw(path):
    rows = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")


def source(path, win=40):
    rows = list(csv.reader(io.StringIO(ncu(["-i", path, "--page", "source", "--csv"]))))
    hi = [i for i, r in enumerate(rows[:10]) if len(r) > 5][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    f = lambda r, n: float(r[ix[n]] or 0) if r[ix[n]].replace(".", "").isdigit() else 0.0
    te, ts = sum(f(r, "Instructions Executed") for r in data), sum(f(r, "# Samples") for r in data)
    print(f"warp-instructions {te:.3e}, samples {ts:.0f}, sass rows {len(data)}")
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: sum(f(r, h) for r in data) for h in st}
    print("stalls:", ", ".join(f"{k[6:]} {100 * v / ts:.0f}%" for k, v in sorted(tot.items(), key=lambda x: -x[1])[:7]))
    for s in range(0, len(data), win):
        seg = data[s:s + win]
        e, sm = sum(f(r, "Instructions Executed") for r in seg), sum(f(r, "# Samples") for r in seg)
        if e / te > 0.02 or sm / ts > 0.02:
            ops = collections.Counter((r[ix["Source"]].split() or [""])[0].split(".")[0] for r in seg)
            top = max(seg, key=lambda r: f(r, "# Samples"))
            print(f"rows {s:5d}+{win}: exec {100 * e / te:5.1f}% samples {100 * sm / ts:5.1f}%  hot: {top[ix['Source']][:60]!r}  ops {dict(ops.most_common(5))}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "source": lambda p, *a: source(p, *map(int, a))}[sys.argv[1]](*sys.argv[2:])
