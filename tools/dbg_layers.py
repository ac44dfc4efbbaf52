"""Debug: per-layer conv times under MLT_DEBUG_FLAGS (results invalid, timing only)."""
import os, sys, json, subprocess
for flags in (0, 1, 2, 6, 7):
    env = dict(os.environ, MLT_DEBUG_FLAGS=str(flags))
    out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--no-cpu-baseline"], env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        r = j["roofline"]
        print(f"flags={flags}: step {j['ms_per_step']:.2f} ms, convs {r['kernel_ms_per_step']:.2f} ms, per-layer", [round(x, 3) for x in r["per_layer_ms"]])
    except Exception as e:
        print("flags", flags, "failed", e, out.stderr[-500:])
