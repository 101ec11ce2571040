"""Times sampling calls (CUDA events, device-resident inputs) for several step counts: separates the per-call fixed cost (encoder,
tables) from the per-step cost.  Env toggles (MDTB200_BRANCHES, MDTB200_CROSS_FUSED, ...) are read at handle creation."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from mdt_policy_b200 import gc_sampling as gcs
from mdt_policy_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = H.build_product(H.mdtv_inner_cfg(4, 4, precision="bf16x3"), 3, "trained")
inp = {k: v.cuda() for k, v in synthetic_inputs(B, seed=4).items()}
state = {"state_images": inp["state_images"], "modality": "lang"}
res = {}
for n in (1, 2, 5, 10, 20):
    sig = gcs.get_sigmas_exponential(n, 0.001, 80.0, "cuda")
    for _ in range(3):
        gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True)
    e1.record(); torch.cuda.synchronize()
    res[n] = e0.elapsed_time(e1) / reps
per_step = (res[20] - res[10]) / 10
print({k: round(v, 3) for k, v in res.items()}, "ms; per-step %.1f us; fixed %.1f us" % (per_step * 1e3, (res[10] - 10 * per_step) * 1e3),
      {k: os.environ.get(k) for k in ("MDTB200_BRANCHES", "MDTB200_CROSS_FUSED", "MDTB200_PDL", "MDTB200_BN_D")})
