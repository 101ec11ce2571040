"""Golden vectors of the reference PerceiverResampler (authoring container only):  python tests/golden/make_golden_perceiver.py

The reference module imports `einops_exts.rearrange_many` (not installed, not vendored; call site perceiver_resampler.py:66):
a two-line stand-in with the published semantics (apply einops.rearrange to every tensor) is injected at import time.
Shipped MDT-V configuration (conf/model/mdtv_agent.yaml:28-32): dim 384, depth 6, 8 heads x 64, 1 time embedding, 3 latents."""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mdt_policy_b200.synthetic import synthetic_tensor  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.golden.make_golden import save  # noqa: E402


@torch.no_grad()
def main():
    ref_shim.load_reference()
    import einops
    ex = types.ModuleType("einops_exts")
    ex.rearrange_many = lambda tensors, pattern, **kw: [einops.rearrange(t, pattern, **kw) for t in tensors]
    sys.modules["einops_exts"] = ex
    from mdt.models.networks.transformers.perceiver_resampler import PerceiverResampler
    out = {}
    for tag, (depth, n_lat, B, nf) in {"shipped": (6, 3, 5, 392), "small": (2, 5, 3, 40), "shipped_init": (6, 3, 5, 392)}.items():
        m = PerceiverResampler(dim=384, depth=depth, dim_head=64, heads=8, num_time_embeds=1, num_latents=n_lat).eval()
        named = [(n, tuple(p.shape)) for n, p in m.named_parameters()]
        assert named == H.perceiver_shapes(depth, n_lat), "tests/helpers.perceiver_shapes out of sync with the reference"
        m.load_state_dict(H.perceiver_state(named, 51, "init" if tag.endswith("_init") else "trained"))
        x = synthetic_tensor(f"perceiver.x.{tag}", (B, 1, nf, 384), 52, "init") * 50.0
        out[f"out_{tag}"] = m(x)
        out[f"out64_{tag}"] = m.double()(x.double()).float()          # the reference in fp64: how much of an error is fp32 noise
        if tag == "small":
            out["names"] = torch.zeros(1)
    save("perceiver", meta=dict(case="perceiver", weight_seed=51, input_seed=52, shipped=[6, 3, 5, 392], small=[2, 5, 3, 40]), **out)


if __name__ == "__main__":
    main()
