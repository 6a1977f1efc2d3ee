"""Developer A/B harness: build variants of libb2of.so with extra -D macros (here, on CPU), then on the GPU box
run a Farneback parity check against live cv2 and a short device-resident bench for each.

  python scripts/ab_variants.py build name1:DEF1,DEF2 name2:DEF3 ...     (here)
  python scripts/ab_variants.py run name1 name2 ...                      (under gpurun; writes gpurun_out/ab_<name>.json)
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "hackathonopticalflow_b200", "csrc", "variants")

def vpath(name):
    return os.path.join(VDIR, f"libb2of_{name}.so")

CHECK = r'''
import sys, json, numpy as np, cv2, torch
sys.path.insert(0, %r)
from hackathonopticalflow_b200 import cv2compat as b2, synth
out = {}
for (h, w) in [(1080, 1920), (270, 480), (135, 241), (37, 53)]:
    fr = synth.sequence(h, w, 2, seed=1000)
    ref = cv2.calcOpticalFlowFarneback(fr[0], fr[1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    mine = b2.calcOpticalFlowFarneback(fr[0], fr[1], None, 0.5, 3, 15, 3, 5, 1.2, 0)
    d = np.sqrt(((ref - mine) ** 2).sum(-1))
    out[f"{h}x{w}"] = [float(d.mean()), float(d.max())]
print("PARITY", json.dumps(out))
''' % ROOT

def main():
    mode = sys.argv[1]
    if mode == "build":
        from hackathonopticalflow_b200 import _lib
        os.makedirs(VDIR, exist_ok=True)
        for spec in sys.argv[2:]:
            name, _, defs = spec.partition(":")
            _lib.build(defines=[d for d in defs.split(",") if d], out=vpath(name))
            print("built", vpath(name))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    steps = os.environ.get("AB_STEPS", "8")
    for name in sys.argv[2:]:
        env = dict(os.environ)
        if name != "default":
            env["B2OF_LIB"] = vpath(name)
        try:
            r1 = subprocess.run([sys.executable, "-c", CHECK], env=env, capture_output=True, text=True, timeout=180)
        except subprocess.TimeoutExpired:
            print(json.dumps({"name": name, "err": "parity check timed out (hang?)"}))
            continue
        par = [l for l in r1.stdout.splitlines() if l.startswith("PARITY")]
        r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "3", "--no-cpu",
                             "--no-e2e", "--no-extras"], env=env, capture_output=True, text=True, timeout=300)
        line = None
        for l in r2.stdout.splitlines():
            if l.startswith("{"):
                line = json.loads(l)
        res = {"name": name, "parity": json.loads(par[0][7:]) if par else r1.stderr[-800:],
               "value": line and line["value"], "kernel_ms": line and line["kernel_ms_per_step"],
               "roof": line and line["roofline"]["frac"], "err": None if line else r2.stderr[-800:]}
        with open(os.path.join(ROOT, "gpurun_out", f"ab_{name}.json"), "w") as f:
            json.dump(res, f)
        print(json.dumps(res))

if __name__ == "__main__":
    main()
