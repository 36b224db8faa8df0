"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md):
UTCHMMA = tcgen05.mma, UTMALDG = TMA load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, HMMA = mma.sync.
usage: python tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sl_hwgat_b200", "lib", "libhwgat_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
WANT = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "HMMA", "MUFU.EX2", "FFMA", "ATOMG", "RED"]
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m:
        op = m.group(1)
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[cur][w] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
print("# SASS instruction summary of libhwgat_b200.so (sm_100a)\n")
print("`cuobjdump -sass sl_hwgat_b200/lib/libhwgat_b200.so`, static instruction counts per kernel "
      "(`tools/sass_summary.py`).  UTCHMMA = `tcgen05.mma` (`.2CTA` = `cta_group::2`), UTMALDG = TMA tensor load, "
      "UBLKCP = bulk copy, LDTM / STTM = `tcgen05.ld` / `tcgen05.st`, UTCBAR = `tcgen05.commit`, HMMA = `mma.sync`.\n")
print("| kernel | " + " | ".join(WANT) + " |")
print("|---|" + "---|" * len(WANT))
tot = collections.Counter()
for fn, c in counts.items():
    name = demangle(fn)
    name = re.sub(r"\(.*", "", name).replace("hwgat::", "")
    name = re.sub(r"^void ", "", name)
    print(f"| `{name}` | " + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
    tot.update(c)
print("| **total** | " + " | ".join(str(tot[w]) for w in WANT) + " |")
