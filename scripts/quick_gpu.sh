#!/bin/bash
# quick GPU regression: parity tests + bench summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_quick.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.log").read().strip().splitlines()[-1])
print({k: round(v,3) for k,v in d["kernel_ms_per_step"].items()}, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "roof", round(d["roofline"]["frac"],3))
PY
