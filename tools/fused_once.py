"""Runs a few B=256 sampling calls through the fused decoder (for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from mdt_policy_b200 import gc_sampling as gcs
from mdt_policy_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
model = H.build_product(H.mdtv_inner_cfg(4, 4, precision="bf16x3"), 3, "trained")
inp = {k: v.cuda() for k, v in synthetic_inputs(B, seed=4).items()}
state = {"state_images": inp["state_images"], "modality": "lang"}
sig = gcs.get_sigmas_exponential(n_steps, 0.001, 80.0, "cuda")
for _ in range(4):
    out = gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True)
torch.cuda.synchronize()
print("ok", float(out.abs().max()))
