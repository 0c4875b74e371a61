"""cuobjdump -sass of the built library: counts of the tensor / TMA / TMEM mnemonics per kernel (evidence for profiles/)."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "generalised-gaussian-processes_b200/libggp_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pat = re.compile(r"\b(UTCIMMA|UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|DMMA|HMMA|IMMA|LDGSTS|SYNCS)\b")
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        for k in pat.findall(line): counts[cur][k] += 1
print("# cuobjdump -sass %s: tensor / TMA / TMEM mnemonics per kernel" % so.split("/")[-1])
tot = collections.Counter()
for k, c in counts.items():
    if c:
        print(demangle(k)[:150]); print("    ", dict(c)); tot.update(c)
print("# total", dict(tot))
