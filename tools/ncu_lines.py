"""Per-source-line executed instructions / stall samples from an .ncu-rep (needs -lineinfo).
usage: ncu_lines.py report [topn] [kernel-substring]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = sys.argv[3] if len(sys.argv) > 3 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
allrows = []
fname = "?"; hdr = None; func = "?"; seen = set(); skip = False
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": fname = row[1].split("/")[-1]; continue
    if row[0] == "Function Name":
        func = row[1]
        continue
    if row[0] == "Kernel Name":
        func = row[1]; continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or row[0] == "": continue
    if kern and kern not in func: continue
    try:
        n = int(row[hdr.index("Instructions Executed")]); s = int(row[hdr.index("# Samples")])
    except (ValueError, IndexError):
        continue
    if n or s: allrows.append((s, n, fname, row[0], row[1].strip()[:105], func[:30]))
tot = sum(r[1] for r in allrows) or 1; ts = sum(r[0] for r in allrows) or 1
print("total warp-instr", tot, "samples", ts)
for s, n, f, ln, src, fn in sorted(allrows, reverse=True)[:topn]:
    print("%5.2f%% samp %5.2f%% instr  %s:%s  %s" % (100.0 * s / ts, 100.0 * n / tot, f, ln, src))
