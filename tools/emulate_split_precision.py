"""CPU emulation of the tensor-core operand precisions (decides the GEMM mode before spending GPU time):
every big GEMM of the oracle is replaced by a split-operand product with fp32 accumulation, then the 10-step DDIM
result is compared with the reference's fp32 / fp64 goldens (config 2, first 64 samples)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mdt_oracle as orc          # noqa: E402
from tests import helpers as H                # noqa: E402
from mdt_policy_b200.synthetic import synthetic_inputs  # noqa: E402

_orig = F.linear


def split(x, dt):
    hi = x.to(dt).float()
    lo = (x - hi).to(dt).float()
    return hi, lo


def make_linear(mode):
    def lin(x, w, b=None):
        if w.shape[0] < 64 or w.shape[1] < 64 or mode == "fp32":      # action_emb / action_pred stay fp32
            return _orig(x, w, b)
        dt = torch.bfloat16 if "bf16" in mode else torch.float16
        xh, xl = split(x, dt)
        wh, wl = split(w, dt)
        if mode.endswith("x1"):
            out = _orig(xh, wh)
        elif mode.endswith("x3"):
            out = _orig(xh, wh) + (_orig(xl, wh) + _orig(xh, wl))
        elif mode.endswith("x4"):
            out = _orig(xh, wh) + (_orig(xl, wh) + _orig(xh, wl) + _orig(xl, wl))
        return out if b is None else out + b
    return lin


def main():
    for profile in ("trained", "init"):
        meta, gold = H.load_golden(f"config2_{profile}")
        P = H.oracle_params(H.mdtv_shapes(4, 4), meta["weight_seed"], profile)
        cfg = orc.OracleCfg()
        inp = synthetic_inputs(256, seed=meta["input_seed"])
        n = 64
        st = {"state_images": inp["state_images"][:n], "modality": "lang"}
        for smin in (0.001, 1.0):
            sig = orc.get_sigmas_exponential(10, smin, 80.0)
            r32, r64 = gold[f"ddim_smin{smin}_f32"][:n], gold[f"ddim_smin{smin}_f64"][:n].float()
            for mode in ("fp32", "bf16x1", "bf16x3", "bf16x4", "fp16x3"):
                F.linear = make_linear(mode)
                try:
                    out = orc.sample(P, cfg, st, inp["x_T"][:n], inp["goal"][:n], sig, "ddim")
                finally:
                    F.linear = _orig
                print(f"{profile:8s} smin={smin:<6} {mode:7s} max|x-ref32|={float((out - r32).abs().max()):.2e} "
                      f"max|x-ref64|={float((out - r64).abs().max()):.2e}  (|ref32-ref64|={float((r32 - r64).abs().max()):.2e})")


if __name__ == "__main__":
    main()
