"""Stall-sample summary of one kernel from an ncu report: python tools/ncu_stalls.py <report.ncu-rep> <kernel-regex> [bin-bytes]
Reads `ncu --page source --csv` (SASS view) and prints the stall reasons of the whole kernel, a histogram over code regions
(`bin-bytes` of SASS each: the roles of a warp-specialised kernel are contiguous regions) and the instructions with most samples."""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
binb = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[hdr_i - 1][1] if hdr_i else kern)
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hdr_i + 1:]:          # first view of the first matching kernel only (the page repeats the header per view)
    if r and r[0] == "Address":
        break
    if len(r) == len(hdr):
        data.append(r)
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def f(r, k):
    try:
        return float(r[ix[k]])
    except ValueError:
        return 0.0


tot = sum(f(r, "# Samples") for r in data)
inst = sum(f(r, "Instructions Executed") for r in data)
print("samples %d, warp instructions %.1f M" % (tot, inst / 1e6))
print("stall reasons (share of all samples):")
for k in sorted(keys, key=lambda k: -sum(f(r, k) for r in data)):
    v = sum(f(r, k) for r in data)
    if v > 0.005 * tot:
        print("  %-22s %5.1f %%" % (k, 100 * v / tot))
bins = collections.OrderedDict()
for r in data:
    b = int(r[ix["Address"]], 16) // binb
    e = bins.setdefault(b, [0.0, 0.0] + [0.0] * len(keys))
    e[0] += f(r, "# Samples"); e[1] += f(r, "Instructions Executed")
    for i, k in enumerate(keys):
        e[2 + i] += f(r, k)
print("code regions of %d bytes (samples, M warp instructions, top two stall reasons):" % binb)
base = min(bins)
for b, e in bins.items():
    if e[0] > 0.003 * tot:
        top = sorted(range(len(keys)), key=lambda i: -e[2 + i])[:2]
        print("  +0x%05x %7d %8.1f   %s" % ((b - base) * binb, e[0], e[1] / 1e6,
                                           ", ".join("%s %.0f%%" % (keys[i][6:], 100 * e[2 + i] / max(e[0], 1)) for i in top)))
print("instructions with most samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:25]:
    st = {k: f(r, k) for k in keys}
    print("  %6d %-18s %s" % (f(r, "# Samples"), max(st, key=st.get)[6:], r[ix["Source"]][:100]))
