"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library: tcgen05 MMA (UTCHMMA), tensor-memory loads (LDTM),
TMA tensor loads / stores (UTMALDG / UTMASTG), bulk copies (UBLKCP), reductions / atomics (REDG / ATOMG).
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "event_flow_b200/lib/libeventflow.so"
MN = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "REDG", "ATOMG", "FFMA2")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for k in MN:
        if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
            counts[name][k] += 1
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print(f"cuobjdump -sass {LIB}: instruction counts per kernel (kernels without any of these mnemonics omitted)")
print(f"{'kernel':78s}" + "".join(f"{k:>9s}" for k in MN))
for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
    c = counts[n]
    if sum(c.values()) == 0:
        continue
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{d[:78]:78s}" + "".join(f"{c[k]:9d}" for k in MN))
