"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UBLKCP / UTMALDG / UTMASTG = bulk-copy engine (TMA), HMMA = legacy mma.sync (must be 0).
Usage: python tools/sass_evidence.py > profiles/r1_sass_evidence.txt   (cross-compiled library, no GPU needed)"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "msra_practice_project_b200", "libb2r.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|LDTM|STTM|UBLKCP|UTMALDG|UTMASTG|UTCBAR|HMMA|HGMMA|MUFU\.SIN|MUFU\.COS|MUFU\.EX2|SYNCS)\b")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        for t in pat.findall(line):
            counts[cur][t] += 1
dem = subprocess.run(["cu++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS mnemonic counts per kernel of libb2r.so (sm_100a), cuobjdump -sass; HMMA / HGMMA must be absent")
for name, (k, c) in zip(dem, counts.items()):
    if not c:
        continue
    short = name.replace("b2r::", "").replace("tc::", "").replace("(int)", "").replace("(bool)", "").replace("void ", "")
    short = re.sub(r"\((const|b2r|float|long|int|unsigned).*", "", short)
    print(f"{short[:44]:44s} " + "  ".join(f"{t}={n}" for t, n in sorted(c.items())))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("total:", dict(tot))
