"""Per-source-line summary of one kernel from an `ncu --set full --import-source on` report (run where ncu can read it).
usage: python profiles/src_hot_lines.py report.ncu-rep <kernel regex> [top]"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0]:
        lines.append((fname, r))
if not lines:
    sys.exit("no source lines")
ix = {n: i for i, n in enumerate(hdr)}
S, I, W = ix["# Samples"], ix["Instructions Executed"], ix["L1 Wavefronts Shared"]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(int(r[S]) for _, r in lines)
tot_i = sum(int(r[I]) for _, r in lines)
tot_w = sum(int(r[W] or 0) for _, r in lines)
print(f"# total samples {tot_s}, warp instructions {tot_i}, shared wavefronts {tot_w}")
lines.sort(key=lambda fr: -int(fr[1][S]))
for f, r in lines[:top]:
    st = sorted(((int(r[ix[n]]), n[6:]) for n in stalls), reverse=True)[:3]
    sts = " ".join(f"{n}:{v}" for v, n in st if v)
    print(f"{100*int(r[S])/tot_s:5.1f}% smp {100*int(r[I])/tot_i:5.1f}% ins {100*int(r[W] or 0)/max(tot_w,1):5.1f}% wav  {f}:{r[0]:>4}  {r[1][:90]}  [{sts}]")
