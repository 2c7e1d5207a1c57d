import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["clocks"]["sm_mhz"], {k:v["ms_per_launch"] for k,v in d["roofline"]["families"].items()}, d["roofline"]["other_ms_per_step"]["layernorm"])
