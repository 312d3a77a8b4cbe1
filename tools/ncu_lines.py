"""Per-CUDA-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on): python tools/ncu_lines.py rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
out = []; fname = ""; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] in ("", "Function Name"): continue
    si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
    try: out.append((int(r[si] or 0), int(r[ie] or 0), fname, r[0], r[1].strip()))
    except ValueError: pass
tot = sum(o[0] for o in out); toti = sum(o[1] for o in out)
print("samples", tot, "warp instructions", toti)
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%6d %5.1f%% %11d %5.1f%%  %s:%s  %s" % (s, 100.0 * s / tot, i, 100.0 * i / max(toti, 1), f, ln, src[:100]))
