"""CPU emulation of the numerics of the tcgen05 CU path (fp16 operands, fp32 accumulate, fp16 activation storage) to rank
error sources against the fp32 oracle without a GPU.  usage: python tools/emulate_cu_precision.py [size] [n]"""
import os, sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastintercu_vvc_b200 import pack_weights as pw
from oracle import ref_arch

LEVELS = ((0, 2), (2, 5), (5, 9), (9, 15))
def softmax_levels(lg):
    out = np.empty_like(lg)
    for a, b in LEVELS:
        e = np.exp(lg[:, a:b] - lg[:, a:b].max(1, keepdims=True)); out[:, a:b] = e / e.sum(1, keepdims=True)
    return out

def h16(t): return t.to(torch.float16).to(torch.float64)

def emulate(sd, size, x, pocqp, pool_fp32, act_round=True, w_round=True, sc_exact=True):
    sd = pw.normalise_state_dict(sd)
    q = (lambda w: torch.from_numpy(pw.quantize_fp16_diffused(w).astype(np.float64))) if w_round else (lambda w: torch.from_numpy(w.astype(np.float64)))
    rnd = h16 if act_round else (lambda t: t)
    y = rnd(F.conv2d(torch.from_numpy(x.astype(np.float64)), torch.from_numpy(sd["conv1.weight"].astype(np.float64)), padding=1))
    feats = []
    for L in range(5):
        for b in range(2):
            p = f"layer{L}.{b}"
            w1, b1 = pw.fold_bn(sd[f"{p}.conv1.weight"], sd, f"{p}.bn1")
            w2, b2 = pw.fold_bn(sd[f"{p}.conv2.weight"], sd, f"{p}.bn2")
            stride = 2 if b == 0 else 1
            t = rnd(F.relu(F.conv2d(y, q(w1), torch.from_numpy(b1.astype(np.float64)), stride=stride, padding=1)))
            o = F.conv2d(t, q(w2), torch.from_numpy(b2.astype(np.float64)), padding=1)
            if b == 0:
                ws, bs = pw.fold_bn(sd[f"{p}.shortcut.0.weight"], sd, f"{p}.shortcut.1")
                wsq = torch.from_numpy(ws.astype(np.float64)) if sc_exact else q(ws)
                o = o + F.conv2d(y, wsq, torch.from_numpy(bs.astype(np.float64)), stride=stride)
            else:
                o = o + y
            o = F.relu(o)
            y = rnd(o)
        if L >= 1:
            feats.append((o if pool_fp32 else y).mean((2, 3)))
    lg = []
    pq = torch.from_numpy(pocqp.astype(np.float64))
    for i, f in enumerate(feats):
        w = torch.from_numpy(sd[f"branch{i+1}.weight"].astype(np.float64)); bb = torch.from_numpy(sd[f"branch{i+1}.bias"].astype(np.float64))
        lg.append(torch.cat([f, pq], 1) @ w.T + bb)
    return torch.cat(lg, 1).numpy()

if __name__ == "__main__":
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    torch.set_num_threads(8)
    sd = ref_arch.make_cu_state_dict(10, size)
    orgpred, pocqp = ref_arch.synth_cus(n, size, 10)
    x = ref_arch.stage_numpy(orgpred)
    ref = ref_arch.forward_cu_logits(ref_arch.build_cu_model(sd), x, pocqp).astype(np.float64)
    pr = softmax_levels(ref)
    for name, kw in (("gpu-like (pool fp16)", dict(pool_fp32=False)), ("pool fp32", dict(pool_fp32=True)),
                     ("no act rounding", dict(pool_fp32=True, act_round=False)), ("no weight rounding", dict(pool_fp32=True, w_round=False)),
                     ("exact", dict(pool_fp32=True, act_round=False, w_round=False))):
        lg = emulate(sd, size, x, pocqp, **kw)
        dl = np.abs(lg - ref); dp = np.abs(softmax_levels(lg) - pr)
        print(f"{size:3d} {name:22s} |dlogit| mean {dl.mean():.2e} max {dl.max():.2e} per-level max {[float('%.1e' % dl[:, a:b].max()) for a, b in LEVELS]} | |dprob| max {dp.max():.2e}")
