"""Per-region stall-sample summary of one kernel from an ncu report captured with --import-source on (needs -lineinfo):
    python tools/ncu_regions.py gpurun_out/x.ncu-rep [marker-regex]
Prints the cumulative warp-stall samples at every instruction matching the marker regex (barriers, exits, ... by default),
i.e. how the kernel's time splits over the code between those markers, plus the instructions with the most samples."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
marker = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"\bBAR\b|EXIT|WARPSYNC\.ALL")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr, data = None, []
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        tot[s] += int(r[ix[s]] or 0)
total = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("instructions", len(data), "samples", total, [(s[6:], v) for s, v in tot.most_common(8)])
cum = last = 0
for n, r in enumerate(data):
    cum += int(r[ix["# Samples"]] or 0)
    src = r[ix["Source"]].strip()
    if marker.search(src):
        print("%5d cum=%6d (+%5d = %4.1f%%) exec=%9s  %s" % (n, cum, cum - last, 100.0 * (cum - last) / max(total, 1), r[ix["Instructions Executed"]], src[:70]))
        last = cum
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:76],
          [(s[6:], int(r[ix[s]])) for s in stalls if int(r[ix[s]] or 0) > 25])
