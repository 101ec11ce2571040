"""Per-phase timeline of the persistent fused decoder (MDTB200_FUSED_TRACE=1): for the first 128 phases of a sampling call
prints, averaged over CTAs, when the workers entered the phase, how long they waited for the accumulator / computed, and
when they published -- all relative to the previous phase's publish time."""
import os
import sys

os.environ["MDTB200_FUSED_TRACE"] = "1"
os.environ["MDTB200_FUSED"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: E402
from mdt_policy_b200 import gc_sampling as gcs  # noqa: E402
from mdt_policy_b200.synthetic import synthetic_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_dec = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
model = H.build_product(H.mdtv_inner_cfg(4, n_dec, precision="bf16x3"), 3, "trained")
inp = {k: v.cuda() for k, v in synthetic_inputs(B, seed=4).items()}
state = {"state_images": inp["state_images"], "modality": "lang"}
sig = gcs.get_sigmas_exponential(n_steps, 0.001, 80.0, "cuda")
for _ in range(3):
    out = gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True)
torch.cuda.synchronize()
eng = list(model.inner_model._engines.values())[0]
raw = eng.debug_buffer("fused_trace", 128 * 160 * 8 * 2)
grid = ((B + 11) // 12) * 6
t = raw.view(torch.int32).cpu().numpy().view("int64")[:128 * grid * 8].reshape(128, grid, 8).astype("float64")
names = ["EMB+LN"] + ["QKV", "ATTN", "O", "LN3", "Q", "XATT", "O2", "LN2", "FC", "PRJ0", "PRJ1", "RED+LN"] * 16
G = int(os.environ.get("TRACE_GROUP", "0"))
t = t[:, G * 6:(G + 1) * 6, :]          # one group: its 6 CTAs run in lockstep
print(f"group {G} of {grid // 6}; times in us; enter = workers past the group wait (relative to the group's previous publish, max over CTAs);")
print("mid = accumulator ready (GEMM) / compute done (row, attention); end = published.  min..max over the 6 CTAs")
tot = {}
prev = t[0, :, 0].min()
nph = 1 + 12 * n_dec * n_steps
for p in range(min(128, nph)):
    enter, mid, end = t[p, :, 0] - prev, t[p, :, 1] - prev, t[p, :, 2] - prev
    if p < 1 + 12 * 2:
        ex = " ".join(f"{(t[p, 0, k] - prev) / 1e3:6.2f}" if t[p, 0, k] > 0 else "   -  " for k in range(3, 8))
        print(f"{p:3d} {names[p]:7s} enter {enter.min() / 1e3:6.2f}..{enter.max() / 1e3:6.2f}  mid {mid.min() / 1e3:6.2f}..{mid.max() / 1e3:6.2f}  end {end.min() / 1e3:6.2f}..{end.max() / 1e3:6.2f} | cta0 s3..7: {ex}")
    tot.setdefault(names[p], []).append(end.max() / 1e3)
    prev = t[p, :, 2].max()
print("mean us per phase kind:", {k: round(sum(v) / len(v), 2) for k, v in tot.items()})
print("sum per layer:", round(sum(sum(v) / len(v) for k, v in tot.items() if k != "EMB+LN"), 2), "us")
