"""Summarise the SASS page of an .ncu-rep: executed instructions by opcode, hottest instructions, stall totals."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    lines = b.splitlines()
    name = lines[0]
    if kern and kern not in name: continue
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
    tot = sum(int(r["Instructions Executed"] or 0) for r in rows)
    print("==", name[:80], "SASS lines", len(rows), "warp-instr executed", tot)
    byop = collections.Counter(); samp = collections.Counter()
    for r in rows:
        op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
        op = op.split(".")[0]
        byop[op] += int(r["Instructions Executed"] or 0); samp[op] += int(r["# Samples"] or 0)
    ts = sum(samp.values()) or 1
    for op, n in byop.most_common(22):
        print("   %-10s %6.2f%% of instr   %6.2f%% of samples" % (op, 100.0 * n / tot, 100.0 * samp[op] / ts))
    stalls = collections.Counter()
    for r in rows:
        for k, v in r.items():
            if k.startswith("stall_") and "Not Issued" not in k and v: stalls[k] += int(v)
    print("   stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / ts) for k, v in stalls.most_common(8)))
    break
