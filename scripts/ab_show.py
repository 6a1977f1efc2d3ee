import json, sys, glob, os
names = sys.argv[1:] or [os.path.basename(f)[3:-5] for f in sorted(glob.glob("gpurun_out/ab_*.json"), key=os.path.getmtime)]
for n in names:
    d = json.load(open(f"gpurun_out/ab_{n}.json"))
    if not d.get("value"):
        print(n, "FAILED", str(d.get("err"))[-300:], str(d.get("parity"))[-300:]); continue
    k = d["kernel_ms"]
    par = d["parity"]["1080x1920"] if isinstance(d["parity"], dict) else d["parity"][-200:]
    print(f"{n:14s} {d['value']:8.1f} pairs/s  finest {k['fb_iter_finest']:.3f}  coarse {k['fb_iter_coarse']:.3f}  level {k['fb_level_hpass']:.3f}  poly {k['fb_polyexp']:.3f}  roof {d['roof']:.3f}  parity1080 {par}")
