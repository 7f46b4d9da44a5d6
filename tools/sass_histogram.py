"""Instruction histogram per kernel of the in-tree library (runs here, no GPU): `cuobjdump -sass`
of libf16_b200.so, the mnemonics that tell which hardware paths a kernel uses.
usage: python tools/sass_histogram.py [lib] > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "flake16_framework_b200", "libf16_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
GROUPS = [("tcgen05.mma", r"^UTC\w*MMA"), ("tcgen05.ld/st (TMEM)", r"^(LDTM|STTM)"), ("tcgen05 alloc/commit", r"^(UTCBAR|UTCATOMSWS|UTCCP)"),
          ("bulk copy (TMA)", r"^(UBLKCP|UTMALDG|UTMASTG)"), ("mbarrier", r"^SYNCS"), ("cp.async", r"^LDGSTS"),
          ("mma.sync", r"^HMMA"), ("FP64 (DFMA/DMUL/DADD)", r"^(DFMA|DMUL|DADD)"), ("FP64 reciprocal seed", r"^MUFU\.RCP64H"),
          ("global load", r"^LDG"), ("global store", r"^STG"), ("shared load", r"^LDS"), ("shared store", r"^STS"),
          ("global atomics", r"^(ATOMG|REDG|RED|ATOM)\b"), ("shared atomics", r"^ATOMS"), ("warp shuffle", r"^SHFL"),
          ("warp vote", r"^VOTE"), ("block barrier", r"^BAR"), ("local (spill) ld/st", r"^(LDL|STL)"), ("float min/max", r"^FMNMX")]
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for name, pat in GROUPS:
            if re.match(pat, op):
                counts[kern][name] += 1
                break
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception:
        return n
print("# SASS instruction histogram of %s (cuobjdump -sass; static counts per kernel)" % os.path.relpath(lib, ROOT))
agg = collections.OrderedDict()
for k, c in counts.items():
    name = re.sub(r"<.*", "", demangle(k))
    a = agg.setdefault(name, [0, 0, collections.Counter()])
    a[0] += 1; a[1] += total[k]; a[2].update(c)
for name, (inst, tot, c) in agg.items():
    print("\n%s   (%d instantiation%s, %d SASS instructions)" % (name, inst, "" if inst == 1 else "s", tot))
    print("   " + "  ".join("%s=%d" % (g, c[g]) for g, _ in GROUPS if c[g]))
