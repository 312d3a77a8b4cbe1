"""Per-kernel SASS mnemonic counts of the in-tree library (evidence that the tcgen05 / TMEM / TMA path is what got built):
    python tools/sass_table.py [libskps_b200.so] > profiles/<round>_sass_mnemonics.txt
UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, HMMA = legacy warp-level mma.sync."""
import collections, os, re, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "peppa_pig_face_landmark_b200", "libskps_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "FFMA", "LDS", "STS", "LDG", "STG"]
cur, cnt = None, collections.defaultdict(collections.Counter)
for l in sass.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1)
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if cur and m:
        cnt[cur]["_n"] += 1
        op = m.group(1).split(".")[0]
        if op in pats:
            cnt[cur][op] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -sass %s (sm_100a), static instruction counts per kernel" % os.path.basename(so))
print("%-60s %6s " % ("kernel", "instr") + " ".join("%7s" % p for p in pats))
tot = collections.Counter()
for n, d in zip(names, dem):
    c = cnt[n]
    tot.update(c)
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("skps::", "")[:60]
    print("%-60s %6d " % (short, c["_n"]) + " ".join("%7d" % c[p] for p in pats))
print("%-60s %6d " % ("TOTAL", tot["_n"]) + " ".join("%7d" % tot[p] for p in pats))
