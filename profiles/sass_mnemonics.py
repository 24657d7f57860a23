"""cuobjdump -sass of the built library: per kernel, how often the Blackwell-specific mnemonics appear
(UBLKCP = cp.async.bulk / TMA non-tensor, SYNCS = mbarrier, LDG.E...256 = 256-bit loads, LTC64B = 64-byte L2 fill hint).
usage: python profiles/sass_mnemonics.py > profiles/r02_sass_tma.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nextpolish2_b200", "libnp2gpu.so")], capture_output=True, text=True).stdout
cur, tab = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        tab[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for pat, name in ((r"\bUBLKCP", "UBLKCP"), (r"\bSYNCS", "SYNCS"), (r"LDG\.E\S*\.256", "LDG.E.256"), (r"LTC64B", "LTC64B"),
                      (r"\bREDUX", "REDUX"), (r"\bUTMA", "UTMA"), (r"\bTCGEN|UTC", "tcgen05")):
        if re.search(pat, line):
            tab[cur][name] += 1
print("cuobjdump -sass nextpolish2_b200/libnp2gpu.so: Blackwell-specific mnemonics per kernel (B200_PROFILING.md)")
print("UBLKCP = cp.async.bulk (TMA, non-tensor), SYNCS = mbarrier, LDG.E.256 = 256-bit load (sm_100+), LTC64B = 64-byte L2 fill\n")
for k, c in tab.items():
    if c:
        print("%-110s %s" % (k[:110], dict(c)))
