"""The cluster chain kernel (convs 1..15 of a 1-2 CTU call as one launch) against the per-layer kernels: bit-identical results and
the latency of the in-encoder call.  Run on the B200 box under a timeout: timeout -s KILL 120 python tools/chain_check.py"""
import os, subprocess, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def child():
    import fastintercu_vvc_b200 as pkg
    from fastintercu_vvc_b200.synth import make_state_dict, synth_ctus
    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    pkg.write_blob(make_state_dict(10), blob)
    ctus, pq = synth_ctus(24, 5)
    with pkg.MltPredictor(blob, max_batch=64) as p:
        full = p.predict_batch_dense(ctus, pq)  # per-layer kernels (n = 24)
        if os.environ.get("CHAIN_CHECK_SHORT"):
            for i in range(6):
                p.predict_ctu(ctus[i, 0], ctus[i, 1], int(pq[i, 0]), int(pq[i, 1]))
            return
        l0 = p.launch_count
        ok = True
        for i in range(24):
            one = p.predict_ctu(ctus[i, 0], ctus[i, 1], int(pq[i, 0]), int(pq[i, 1]))
            ok &= one.tobytes() == full[i].tobytes()
        print("launches per single-CTU call:", (p.launch_count - l0) / 24, " bit-identical to the batch path:", ok)
        two = p.predict_batch_dense(ctus[4:6], pq[4:6])
        print("pair bit-identical:", two.tobytes() == full[4:6].tobytes())
        again = [p.predict_ctu(ctus[3, 0], ctus[3, 1], int(pq[3, 0]), int(pq[3, 1])).tobytes() for _ in range(50)]
        print("deterministic over 50 calls:", len(set(again)) == 1)
        for _ in range(50):
            p.predict_ctu(ctus[0, 0], ctus[0, 1], int(pq[0, 0]), int(pq[0, 1]))
        t0 = time.perf_counter()
        for _ in range(500):
            p.predict_ctu(ctus[0, 0], ctus[0, 1], int(pq[0, 0]), int(pq[0, 1]))
        print(f"single-CTU call: {(time.perf_counter() - t0) / 500 * 1e6:.1f} us")
    os.unlink(blob)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(); sys.exit(0)
    envs = ({"MLT_CHAIN": "1"}, {}, {"MLT_CHAIN": "1", "MLT_CHAIN_TRACE": "1", "CHAIN_CHECK_SHORT": "1"})
    if os.environ.get("CHAIN_TRACE_LAYERS"):
        envs = tuple({"MLT_CHAIN": "1", "MLT_CHAIN_TRACE": "1", "CHAIN_CHECK_SHORT": "1", "MLT_CHAIN_TRACE_LAYER": l} for l in os.environ["CHAIN_TRACE_LAYERS"].split(","))
    for env in envs:
        r = subprocess.run(["timeout", "-s", "KILL", "90", sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True)
        print(env or "per-layer kernels (default)", "rc", r.returncode, "\n", r.stdout.strip(), r.stderr[-400:], flush=True)
