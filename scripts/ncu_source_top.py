"""Per-instruction stall summary of one kernel from an .ncu-rep (read here, no GPU).
usage: python scripts/ncu_source_top.py rep [topN]   -> top instructions by stall samples + per-stall-reason totals"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line: kernel name; second: header
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; body = [r for r in rows[1:] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
S = ci["Warp Stall Sampling (All Samples)"]; X = ci["Instructions Executed"]
tot = sum(int(r[S]) for r in body); totx = sum(int(r[X]) for r in body)
print("kernel:", lines[0][:120]); print("total samples", tot, "warp instr", totx, "n sass", len(body))
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]]) for r in body) for h in reasons}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
print("--- cumulative by address order (idx, samples%, instr%) every 50 sass")
cs = cx = 0
for i, r in enumerate(body):
    cs += int(r[S]); cx += int(r[X])
    if i % 50 == 49 or i == len(body) - 1:
        print(f"  sass[{i-49 if i>=49 else 0}:{i+1}] cum samples {100*cs/tot:.1f}% cum instr {100*cx/totx:.1f}%")
print("--- top instructions")
order = sorted(range(len(body)), key=lambda i: -int(body[i][S]))[:top]
for i in sorted(order):
    r = body[i]
    rs = {h[6:]: int(r[ci[h]]) for h in reasons if int(r[ci[h]])}
    rs = dict(sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print(f"  [{i:4d}] {100*int(r[S])/tot:5.2f}%  x{int(r[X]):>9d}  {r[ci['Source']].strip()[:60]:60s} {rs}")
