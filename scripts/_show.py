import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(l['value'], 1), round(l['ms_per_step'], 3), 'serial', round(l['kernel_timing_pass']['ms_per_step'], 3), 'path', round(l['path_roofline']['frac'], 4), 'kern', round(l['roofline']['frac'], 4), {k: round(v, 3) for k, v in l['kernel_ms_per_step'].items()})
