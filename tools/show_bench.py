import json,sys
d=json.loads(sys.stdin.readline())
print(d["config"].get("dp_exchange"), round(d["value"]), round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k:(round(v["tflops"]),round(v["ms_per_step"],3)) for k,v in d["roofline"]["per_class"].items()}, {k:round(v,3) for k,v in d.get("phases_ms_per_step",{}).items()})
