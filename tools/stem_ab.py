"""A/B of stem kernel variants on the B200 box: device-resident 3840-CTU steps, whole-step time and the stem's own time
(serial per-kernel profile).  usage: python tools/stem_ab.py  (spawns itself once per variant)"""
import os, subprocess, sys, tempfile, time
import numpy as np

VARIANTS = [{"MLT_STEM5_STAGERS": "4"}, {"MLT_STEM5_DBG": "1"}, {"MLT_STEM5_DBG": "2"}, {"MLT_STEM5_DBG": "4"}, {"MLT_STEM5_DBG": "7"},
            {"MLT_STEM_OLD": "1"}]
if os.environ.get("STEM_AB_VARIANTS"):
    VARIANTS = [dict(kv.split("=") for kv in v.split(",")) for v in os.environ["STEM_AB_VARIANTS"].split(";")]

def child(blob):
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fastintercu_vvc_b200 as pkg
    from fastintercu_vvc_b200.capi import RESULT_DTYPE
    import bench
    n = 3840
    orgpred, pocqp = bench.synth_frames(n, 1000)
    d_in, d_pq = torch.from_numpy(orgpred).cuda(), torch.from_numpy(pocqp).cuda()
    d_out = torch.zeros(n * RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    with pkg.MltPredictor(blob, device=0, max_batch=n) as p:
        step = lambda: p.predict_batch_device(n, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), st.cuda_stream)
        for _ in range(5): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(60): step()
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 60
        sus = lat = 0.0
        if os.environ.get("STEM_AB_SUSTAIN"):  # power-capped regime: ~3 s back to back, then the one-CTU latency
            e0.record(st)
            for _ in range(600): step()
            e1.record(st); torch.cuda.synchronize()
            sus = e0.elapsed_time(e1) / 600
            one = lambda: p.predict_batch_device(1, d_in.data_ptr(), d_pq.data_ptr(), d_out.data_ptr(), st.cuda_stream)
            for _ in range(20): one()
            torch.cuda.synchronize()
            e0.record(st)
            for _ in range(200): one()
            e1.record(st); torch.cuda.synchronize()
            lat = e0.elapsed_time(e1) / 200 * 1000
        p.set_profiling(True)
        prof = np.zeros(18)
        for _ in range(5):
            step(); torch.cuda.synchronize(); prof += p.get_profile()
        prof /= 5
    print(f"step {ms:.3f} ms ({n / ms:.0f} k CTU/s)  stem {prof[0]:.3f} ms  convs {prof[2:17].sum():.3f} ms" +
          (f"  sustained {sus:.3f} ms ({n / sus:.0f} k CTU/s)  one CTU {lat:.1f} us" if sus else ""))

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1]); sys.exit(0)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fastintercu_vvc_b200.pack_weights import write_blob
    from fastintercu_vvc_b200.synth import make_state_dict
    blob = tempfile.NamedTemporaryFile(suffix=".mltw", delete=False).name
    write_blob(make_state_dict(10), blob)
    for v in VARIANTS:
        env = dict(os.environ, **v)
        r = subprocess.run(["timeout", "-s", "KILL", "60", sys.executable, __file__, blob], env=env, capture_output=True, text=True)
        print(v, r.stdout.strip() or r.stderr[-300:], flush=True)
    os.unlink(blob)
