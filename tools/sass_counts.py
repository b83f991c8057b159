"""profiles/r02_sass_counts.txt: tensor-core / TMA instruction counts per kernel of the built library (cuobjdump -sass)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ibo_b200", "lib", "libibo_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
MNEM = ["DMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "ELECT", "LDGSTS"]
counts, cur = {}, None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = dict.fromkeys(MNEM, 0); continue
    if cur:
        for k in MNEM:
            if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line): counts[cur][k] += 1
names = list(counts)
dem = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
rows = []
for n, d in zip(names, dem):
    d = re.sub(r"^void ", "", d); d = re.sub(r"\((int|bool)\)", "", d); d = re.sub(r"\(.*$", "", d); d = d.replace("ibo::", "").replace("(anonymous namespace)::", "")
    c = counts[n]
    if any(c.values()): rows.append((d, "  ".join("%s %d" % (k, c[k]) for k in MNEM if c[k])))
rows.sort()
with open(os.path.join(ROOT, "profiles", "r02_sass_counts.txt"), "w") as fh:
    fh.write("# cuobjdump -sass ibo_b200/lib/libibo_b200.so (tools/sass_counts.py): instruction counts per kernel (DMMA = FP64 tensor pipe,\n"
             "# UTCIMMA = tcgen05.mma kind::i8, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKCP = 1-D bulk TMA, ELECT = elect.sync,\n"
             "# LDGSTS = cp.async)\n")
    for d, c in rows: fh.write("%-58s %s\n" % (d, c))
print(len(rows), "kernels")
