"""Error of the CUDA PerceiverResampler against the oracle as a function of depth (shipped shapes)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H
from oracle import perceiver_oracle as po
from mdt_policy_b200.perceiver import PerceiverResampler
from mdt_policy_b200.synthetic import synthetic_tensor
x = synthetic_tensor("perceiver.x.shipped", (5, 1, 392, 384), 52, "init") * 50.0
for depth in (1, 2, 3, 6):
    P = H.perceiver_state(H.perceiver_shapes(depth, 3), 51)
    m = PerceiverResampler(dim=384, depth=depth, dim_head=64, heads=8, num_latents=3, num_time_embeds=1)
    m.load_state_dict(P); m = m.cuda()
    with torch.no_grad():
        want = po.perceiver_forward(P, x, depth)
        want64 = po.perceiver_forward({k: v.double() for k, v in P.items()}, x.double(), depth).float()
    got = m(x.cuda()).cpu()
    e = (got - want).abs()
    print(f"depth {depth}: max err {float(e.max()):.2e} mean {float(e.mean()):.2e} | oracle fp32 vs fp64 {float((want - want64).abs().max()):.2e} | cuda vs fp64 {float((got - want64).abs().max()):.2e}")
