"""Stall/instruction summary of one kernel of an .ncu-rep: python tools/ncu_stalls.py rep.ncu-rep <kernel-id-filter e.g. :::14> [top]"""
import collections, csv, io, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 14
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for n, hi in enumerate(his):
    h = rows[hi]
    end = his[n + 1] - 1 if n + 1 < len(his) else len(rows)
    print("==", rows[hi - 1][1][:110] if hi else "")
    data = [r for r in rows[hi + 1:end] if len(r) == len(h) and r[h.index("# Samples")].isdigit()]
    ie, so, si = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    agg, ops, tot, ti = collections.Counter(), collections.Counter(), 0, 0
    for r in data:
        tot += int(r[si])
        for i in stall:
            if r[i] not in ("", "0"):
                agg[h[i]] += int(r[i])
        t = r[so].strip().split()
        op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
        ops[op] += int(r[ie]); ti += int(r[ie])
    print("samples", tot, [(k, round(100 * v / max(tot, 1), 1)) for k, v in agg.most_common(7)])
    print("warp instr", ti, [(k, round(100 * v / max(ti, 1), 1)) for k, v in ops.most_common(12)])
    for r in sorted(data, key=lambda r: -int(r[si]))[:top]:
        st = sorted(((h[i], int(r[i])) for i in stall if r[i] not in ("", "0")), key=lambda kv: -kv[1])[:2]
        print("%6s %9s %-64s %s" % (r[si], r[ie], r[so].strip()[:64], st))
