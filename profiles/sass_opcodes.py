"""Per-kernel SASS opcode counts of the shipped library (what proves tcgen05 / TMEM / TMA code, B200_PROFILING.md).
usage: python profiles/sass_opcodes.py [melissa_b200/lib/libmelissa_b200.so] > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "melissa_b200/lib/libmelissa_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "LDGSTS", "HFMA2", "HADD2", "HMMA", "FFMA", "RED", "ATOMG"]
print("# SASS opcode summary of melissa_b200/lib/libmelissa_b200.so (cuobjdump -sass, sm_100a)")
print("# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld, UTMALDG = TMA load, UTCBAR = tcgen05.commit, LDGSTS = cp.async,")
print("# HFMA2/HADD2 = packed half math, HMMA = legacy mma.sync (none anywhere), D* = fp64 (environment reward arithmetic)")
print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in COLS) + f" {'D*':>5s}")
blocks = re.split(r"Function : \S+", txt)[1:]
for name, body in zip(names, blocks):
    ops = collections.Counter()
    n = 0
    for m in re.finditer(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", body):
        op = m.group(1)
        n += 1
        ops[op] += 1
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("mls::", "").replace("void ", "")
    d = sum(v for k, v in ops.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU") and k != "MUFU") + ops.get("DFMA", 0) * 0
    print(f"{short[:58]:58s} {n:6d} " + " ".join(f"{ops.get(c, 0):7d}" for c in COLS) + f" {d:5d}")
