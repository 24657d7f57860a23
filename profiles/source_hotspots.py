"""Per-source-line stall samples of every kernel in an `ncu --set full --import-source on` report.
usage: python profiles/source_hotspots.py gpurun_out/r01end_full.ncu-rep [lines per kernel] > profiles/<tag>_source_hotspots.txt"""
import csv
import io
import subprocess
import sys


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ki, di = rows[0].index("Kernel Name"), rows[0].index("gpu__time_duration.sum")
    seen = []
    for r in rows[2:]:
        name = r[ki].split("(")[0]
        if name not in [s[0] for s in seen]:
            seen.append((name, r[di]))
    return seen


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for name, dur in kernels(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", name.replace("void ", "").split("<")[0],
                              "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = next((i for i, r in enumerate(rows) if r and r[0] == "Line No"), None)
        if hdr is None:
            continue
        h = rows[hdr]
        cs, ci, sec = h.index("# Samples"), h.index("Instructions Executed"), h.index("L2 Theoretical Sectors Global")
        stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        lines, agg, fname = [], {}, ""
        for r in rows:  # one section per source file the kernel's code comes from (headers included)
            if r and r[0] == "File Path":
                fname = r[1].rsplit("/", 1)[-1]
            elif len(r) == len(h) and r[0].isdigit():
                st = {h[i][6:]: I(r[i]) for i in stalls if I(r[i]) > 0}
                lines.append(("%s:%s" % (fname, r[0]), r[1].strip(), I(r[cs]), I(r[ci]), I(r[sec]), st))
                for k, v in st.items():
                    agg[k] = agg.get(k, 0) + v
        ts, ti = max(1, sum(l[2] for l in lines)), max(1, sum(l[3] for l in lines))
        tot = max(1, sum(agg.values()))
        print("=== %s   %s ns   warp instructions %d" % (name, dur, ti))
        print("    stalls: " + ", ".join("%s %.0f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
        for l in sorted(lines, key=lambda x: -x[2])[:top]:
            why = ",".join("%s %d" % kv for kv in sorted(l[5].items(), key=lambda kv: -kv[1])[:2])
            print("    %-22s samples %4.1f%%  inst %4.1f%%  sectors %9d  [%s]  %s" % (l[0], 100 * l[2] / ts, 100 * l[3] / ti, l[4], why, l[1][:90]))
        print()


if __name__ == "__main__":
    main()
