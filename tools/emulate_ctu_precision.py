"""CPU emulation (float64) of the numerics of the tcgen05 CTU path -- fp16 operands, fp32-like accumulate, fp16 activation storage --
to rank error sources of the level-3 logits against the exact network without a GPU.
usage: python tools/emulate_ctu_precision.py [n]"""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastintercu_vvc_b200 import pack_weights as pw
from oracle import ref_arch

def h16(t): return t.to(torch.float16).to(torch.float64)
def hilo(t):
    hi = t.to(torch.float16).to(torch.float64)
    return hi + (t - hi).to(torch.float16).to(torch.float64)

def emulate(sd, x, pocqp, round_stage=(True, True, True, True), w_round=(True, True, True, True), stem5=True, hilo_stage=()):
    """round_stage[L]: activations written inside stage L are stored as fp16; hilo_stage: stages that store fp16 hi + lo pairs instead."""
    sd = pw.normalise_state_dict(sd)
    T = lambda a: torch.from_numpy(np.asarray(a, np.float64))
    qd = lambda w: T(pw.quantize_fp16_diffused(w).astype(np.float64))
    xin = T(x)
    c1 = F.conv2d(xin, T(sd["conv1.weight"]), padding=1)  # exact conv1
    y = c1
    feats = []
    for L in range(4):
        rnd = (hilo if L in hilo_stage else h16) if round_stage[L] else (lambda t: t)
        q = qd if w_round[L] else T
        for b in range(2):
            p = f"layer{L}.{b}"
            w1, b1 = pw.fold_bn(sd[f"{p}.conv1.weight"], sd, f"{p}.bn1")
            w2, b2 = pw.fold_bn(sd[f"{p}.conv2.weight"], sd, f"{p}.bn2")
            stride = 2 if b == 0 else 1
            if L == 0 and b == 0 and stem5:
                # composite stem: W5 rounded to fp16 (plain), conv1 never rounded; shortcut input = fp16(conv1 with fp16 weights)
                t = F.conv2d(c1, T(w1), T(b1), stride=2, padding=1)
                W5 = pw.stem5_composite(sd["conv1.weight"], w1)[0]
                if w_round[0]:
                    dW = torch.from_numpy((W5 * 1.0).astype(np.float16).astype(np.float64) - W5)
                    xp = xin
                    t = t + F.conv2d(xp, dW, stride=2, padding=2)  # rounding error of the composite weights (border terms ignored)
                t = rnd(F.relu(t))
                yq = rnd(c1)
            else:
                t = rnd(F.relu(F.conv2d(y, q(w1), T(b1), stride=stride, padding=1)))
                yq = y
            o = F.conv2d(t, q(w2), T(b2), padding=1)
            if b == 0:
                ws, bs = pw.fold_bn(sd[f"{p}.shortcut.0.weight"], sd, f"{p}.shortcut.1")
                o = o + F.conv2d(yq, T(ws), T(bs), stride=stride)  # hi + lo shortcut weights ~ exact
            else:
                o = o + y
            o = F.relu(o)
            y = rnd(o)
        if L >= 1:
            feats.append(o.mean((2, 3)))  # pooled from the fp32 accumulators
    pq = T(pocqp)
    lg = [torch.cat([f, pq], 1) @ T(sd[f"branch{i+1}.weight"]).T + T(sd[f"branch{i+1}.bias"]) for i, f in enumerate(feats)]
    return torch.cat(lg, 1).numpy()

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    torch.set_num_threads(8)
    sd = ref_arch.make_state_dict(10)
    orgpred, pocqp = ref_arch.synth_ctus(n, 10)
    x = ref_arch.stage_numpy(orgpred)
    A, N = (True,) * 4, (False,) * 4
    ref = emulate(sd, x, pocqp, N, N, stem5=False)
    for name, kw in (("gpu-like", dict()),
                     ("no act rounding", dict(round_stage=N)), ("no weight rounding", dict(w_round=N)),
                     ("act rounding only stage 0", dict(round_stage=(True, False, False, False), w_round=N)),
                     ("act rounding only stage 1", dict(round_stage=(False, True, False, False), w_round=N)),
                     ("act rounding only stage 2", dict(round_stage=(False, False, True, False), w_round=N)),
                     ("act rounding only stage 3", dict(round_stage=(False, False, False, True), w_round=N)),
                     ("w rounding only stage 0", dict(round_stage=N, w_round=(True, False, False, False))),
                     ("w rounding only stage 1", dict(round_stage=N, w_round=(False, True, False, False))),
                     ("w rounding only stage 2", dict(round_stage=N, w_round=(False, False, True, False))),
                     ("w rounding only stage 3", dict(round_stage=N, w_round=(False, False, False, True))),
                     ("gpu-like + hilo stage 3", dict(hilo_stage=(3,))),
                     ("gpu-like + hilo stages 2,3", dict(hilo_stage=(2, 3)))):
        lg = emulate(sd, x, pocqp, **kw)
        dl = np.abs(lg - ref)
        print(f"{name:30s} |dlogit| L1 rms {np.sqrt((dl[:, :2]**2).mean()):.2e} L2 rms {np.sqrt((dl[:, 2:5]**2).mean()):.2e} L3 rms {np.sqrt((dl[:, 5:]**2).mean()):.2e} max {dl[:, 5:].max():.2e}", flush=True)
