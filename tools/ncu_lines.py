"""Aggregate an ncu report's source page per CUDA source line: instructions executed and stall samples.
usage: python tools/ncu_lines.py report.ncu-rep kernel-regex [top]"""
import csv, collections, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; agg = collections.defaultdict(lambda: [0, 0, ""]); ii = si = None
stall_cols = {}
stall_tot = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 2 and r[0] == "Line No":
        ii = r.index("Instructions Executed"); si = r.index("# Samples")
        stall_cols = {c: i for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c}
        continue
    if len(r) > 10 and r[2] == "-":
        try: inst = int(r[ii]); samp = int(r[si])
        except Exception: continue
        a = agg[(cur, int(r[0]))]; a[0] += inst; a[1] += samp; a[2] = r[1]
        for c, i in stall_cols.items():
            try: stall_tot[c] += int(r[i])
            except Exception: pass
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[1] for v in agg.values()) or 1
print(f"total inst {tot}  samples {tots}")
print("stalls:", ", ".join(f"{k[6:]} {100*v/sum(stall_tot.values()):.1f}%" for k, v in stall_tot.most_common(10)))
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f:16s}:{l:4d} inst {100*v[0]/tot:5.1f}%  samples {100*v[1]/tots:5.1f}%  {v[2][:100]}")
