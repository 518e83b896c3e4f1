import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d["ms_per_step"],3), round(d["path"]["frac_of_peak"],3), {k:(round(v["ms"],3)) for k,v in d["kernels"].items()})
