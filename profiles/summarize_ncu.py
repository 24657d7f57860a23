"""Turns the ncu artifacts brought back in gpurun_out/ into the small text summaries committed under profiles/.
usage: python profiles/summarize_ncu.py <launches.csv> <raw.csv from `ncu -i rep --page raw --csv`> <out prefix>"""
import collections
import csv
import sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0][:80], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write("# launches %d, total %.1f us\n" % (sum(a[0] for a in agg.values()), tot))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-82s n=%4d %11.1f us %5.1f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sector_hit_rate.pct",
        "dram__sectors_read.sum", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def raw(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none, selected metrics per captured launch\n")
        for r in rows[2:]:
            f.write("--- %s\n" % r[ki].split("(")[0])
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("    %-66s %s %s\n" % (w, r[i], units[i]))


if __name__ == "__main__":
    launches(sys.argv[1], sys.argv[3] + "_launches.txt")
    raw(sys.argv[2], sys.argv[3] + "_full.txt")
