import os, sys, json, subprocess
for fl in (7,15,31):
    env=dict(os.environ, MLT_DEBUG_FLAGS=str(fl), MLT_NO_SLICE="1")
    out=subprocess.run([sys.executable,"bench.py","--no-cpu-baseline","--steps","10","--warmup","3"],env=env,capture_output=True,text=True,timeout=90).stdout.strip().splitlines()[-1]
    d=json.loads(out); r=d["roofline"]
    print("flags",fl,"ms",round(d["ms_per_step"],3),"clk",d["clocks"]["sm_mhz"],[round(x,3) for x in r["per_layer_ms"]])
