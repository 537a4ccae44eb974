"""Per-source-line executed instructions / stall samples from an .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
allrows = []
fname = "?"; hdr = None
for row in csv.reader(io.StringIO(out)):
    if not row: continue
    if row[0] == "File Path": fname = row[1].split("/")[-1]; continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; continue
    if hdr is None or row[0] == "": continue
    try:
        n = int(row[hdr.index("Instructions Executed")]); s = int(row[hdr.index("# Samples")])
    except (ValueError, IndexError):
        continue
    if n: allrows.append((n, s, fname, row[0], row[1].strip()[:105]))
tot = sum(r[0] for r in allrows); ts = sum(r[1] for r in allrows) or 1
print("total warp-instr", tot)
for n, s, f, ln, src in sorted(allrows, reverse=True)[:topn]:
    print("%5.2f%% instr %5.2f%% samp  %s:%s  %s" % (100.0 * n / tot, 100.0 * s / ts, f, ln, src))
