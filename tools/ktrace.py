"""Real (concurrent) kernel timeline of one sampling call: arms mdtb200_debug_ktrace, replays the graph once and prints, per
kernel instance of ONE sub-batch chain, when its first CTA started relative to the previous instance and how long its CTAs ran."""
import ctypes as C, os, sys, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from mdt_policy_b200 import gc_sampling as gcs
from mdt_policy_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = H.build_product(H.mdtv_inner_cfg(4, 4, precision="bf16x3"), 3, "trained")
inp = {k: v.cuda() for k, v in synthetic_inputs(B, seed=4).items()}
state = {"state_images": inp["state_images"], "modality": "lang"}
sig = gcs.get_sigmas_exponential(n_steps, 0.001, 80.0, "cuda")
for _ in range(3):
    gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True)
torch.cuda.synchronize()
eng = list(model.inner_model._engines.values())[0]
cap = 400000
eng.lib.mdtb200_debug_ktrace(eng.handle, cap, None, 0)
gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True)
torch.cuda.synchronize()
buf = np.zeros(cap * 2, dtype=np.uint64)
n = eng.lib.mdtb200_debug_ktrace(eng.handle, 0, buf.ctypes.data_as(C.c_void_p), cap)
rec = buf[:2 * n].reshape(n, 2)
t = rec[:, 0].astype(np.int64); info = rec[:, 1]
tag = (info >> np.uint64(56)).astype(int); ev = ((info >> np.uint64(52)) & np.uint64(15)).astype(int)
sm = ((info >> np.uint64(40)) & np.uint64(0xfff)).astype(int); grid = ((info >> np.uint64(20)) & np.uint64(0xfffff)).astype(int)
blk = (info & np.uint64(0xfffff)).astype(int)
t0 = t.min(); t = (t - t0) / 1e3
names = {0: "other", 1: "gemm", 2: "ln", 3: "attn", 4: "head", 5: "embed", 6: "cross", 7: "pack", 8: "skinny", 9: "fused"}
print(f"{n} records, span {t.max():.1f} us, distinct SMs {len(set(sm))}")
# kernel instances = groups of start events with the same (tag, grid) arriving as `grid` records; chains interleave, so bucket
# start records per (tag, grid) in time order and cut every `grid` records
starts = collections.defaultdict(list); ends = collections.defaultdict(list)
order = np.argsort(t, kind="stable")
for i in order:
    (starts if ev[i] == 0 else ends)[(tag[i], grid[i])].append(t[i])
inst = []
for key, ts in starts.items():
    g = key[1]
    es = ends.get(key, [])
    for k in range(0, len(ts), g):
        chunk = ts[k:k + g]
        e = es[k:k + g]
        inst.append((chunk[0], chunk[-1], max(e) if e else float("nan"), key))
inst.sort()
print("SM-occupancy: mean concurrent CTAs (by start->end of gemm/cross only) n/a; listing first 70 kernel instances:")
prev = 0.0
for a, b, e, key in inst[:70]:
    print(f"  t={a:8.2f} (+{a - prev:6.2f})  last CTA start +{b - a:5.2f}  end +{e - a:6.2f}  {names[key[0]]:6s} grid={key[1]}")
    prev = a
# per-kernel-kind statistics of duration (first start -> last end) and of the gap to the next instance start
agg = collections.defaultdict(list)
for a, b, e, key in inst:
    if e == e:
        agg[(names[key[0]], key[1])].append(e - a)
for k, v in sorted(agg.items()):
    print(f"{k}: n={len(v)} mean duration {np.mean(v):6.2f} us")
