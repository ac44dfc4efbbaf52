"""Probe the CU path at (size, cap, n) combinations, each in its own process (a device fault kills the context).
usage: python tools/cu_probe.py            -> runs the matrix
       python tools/cu_probe.py size cap n -> one case (for compute-sanitizer)"""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def one(size, cap, n):
    import fastintercu_vvc_b200 as pkg
    from fastintercu_vvc_b200.synth import make_cu_state_dict, synth_cus
    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    pkg.write_cu_blob(make_cu_state_dict(10, size), size, blob)
    base, pq = synth_cus(64, size, 3)
    idx = np.arange(n) % 64
    with pkg.MltCuPredictor(blob, size, max_batch=cap) as p:
        r = p.predict_batch_dense(np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx]))
        r2 = p.predict_batch_dense(np.ascontiguousarray(base[idx]), np.ascontiguousarray(pq[idx]))
    assert r.tobytes() == r2.tobytes()
    assert r[:64].tobytes() == r[64:128].tobytes() if n >= 128 else True
    print("ok", size, cap, n, r["logits"][0, :3])

if __name__ == "__main__":
    if len(sys.argv) == 4:
        one(*map(int, sys.argv[1:]))
    else:
        for size, cap, n in ((64, 300, 256), (64, 1024, 1024), (64, 3840, 256), (64, 2048, 2048), (64, 3840, 3840), (32, 15360, 15360), (16, 61440, 61440)):
            r = subprocess.run([sys.executable, __file__, str(size), str(cap), str(n)], capture_output=True, text=True)
            print(size, cap, n, "rc", r.returncode, (r.stdout.strip() or r.stderr.strip()[-300:]).replace("\n", " | "), flush=True)
