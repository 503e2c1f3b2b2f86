"""Hot instructions of a kernel from `ncu --page source --csv` output: address offset, SASS, samples, executed count,
dominant stall reasons.  usage: python tools/ncu_source_hot.py report.ncu-rep [min_samples]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
ia, isrc, isamp, iexe = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = int(rows[2][ia], 16)
tot = sum(int(r[isamp]) for r in rows[2:] if len(r) > isamp and r[isamp].isdigit())
print(f"# {rows[0][1][:100]}  total samples {tot}")
for r in rows[2:]:
    if len(r) <= isamp or not r[isamp].isdigit():
        continue
    s = int(r[isamp])
    if s < min_s:
        continue
    st = sorted(((int(r[i]), h[6:]) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
    print(f"{int(r[ia], 16) - base:6x} {s:7d} {100.0 * s / tot:5.1f}% exe={r[iexe]:>9s}  {r[isrc].strip()[:70]:70s} {' '.join(f'{h}:{v}' for v, h in st)}")
