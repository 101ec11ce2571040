"""GPU diagnostic (not a test): stage-by-stage error of the CUDA path against the oracle, written as JSON so one
gpurun call tells which stage is wrong.  Usage: python tools/diag.py [precision] > gpurun_out/diag.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mdt_oracle as orc                                     # noqa: E402
from mdt_policy_b200 import GCDenoiser, gc_sampling as gcs              # noqa: E402
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs  # noqa: E402
from bench import inner_cfg                                              # noqa: E402
import torch.nn.functional as F                                          # noqa: E402


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    res = {"precision": precision}
    dev = torch.device("cuda", 0)
    for enc, dec, B in ((0, 1, 4), (1, 1, 4), (2, 2, 4), (4, 4, 256)):
        tag = f"e{enc}d{dec}B{B}"
        try:
            model = GCDenoiser(inner_cfg(enc, dec, precision, B), sigma_data=0.5)
            sd = synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], 3, "trained")
            model.load_state_dict(sd)
            model = model.to(dev).eval()
            cfg = orc.OracleCfg(n_enc_layers=enc, n_dec_layers=dec)
            inp = synthetic_inputs(B, seed=4)
            st = {"state_images": inp["state_images"], "modality": "lang"}
            std = {"state_images": inp["state_images"].to(dev), "modality": "lang"}
            sig = torch.exp(torch.linspace(4.0, -6.0, B))
            x = inp["noise"] * sig[:, None, None]
            with torch.no_grad():
                ctx_w = orc.encode(sd, cfg, st, inp["goal"])
                ctx_g = model.forward_context_only(std, x.to(dev), inp["goal"].to(dev), sig.to(dev)).cpu()
                res[tag + "_ctx"] = float((ctx_g - ctx_w).abs().max())
                raw_w = orc.inner_forward(sd, cfg, st, x, inp["goal"], sig)
                raw_g = model.inner_model(std, x.to(dev), inp["goal"].to(dev), sig.to(dev)).cpu()
                res[tag + "_raw"] = float((raw_g - raw_w).abs().max())
                res[tag + "_raw_scale"] = float(raw_w.abs().max())
                # AdaLN table rows (per-sample sigma): mod = Lin(SiLU(sigma_emb))
                eng = list(model.inner_model._engines.values())[0]
                d = 384
                mod_g = eng.debug_buffer("mod", B * dec * 6 * d).cpu().view(B, dec, 6 * d)
                c = orc.sigma_embedding(sd, cfg, sig)[:, 0]
                for l in range(dec):
                    p = f"inner_model.decoder.blocks.{l}.adaLN_zero.modulation.1."
                    mw = F.linear(F.silu(c), sd[p + "weight"], sd[p + "bias"])
                    res[f"{tag}_mod{l}"] = float((mod_g[:, l] - mw).abs().max())
                den_w = orc.denoiser_forward(sd, cfg, st, x, inp["goal"], sig)
                den_g = model(std, x.to(dev), inp["goal"].to(dev), sig.to(dev)).cpu()
                res[tag + "_den"] = float((den_g - den_w).abs().max())
            sg = gcs.get_sigmas_exponential(10, 0.001, 80.0)
            n = min(B, 16)
            for name in ("ddim", "euler", "heun", "dpmpp_2m"):
                w = orc.sample(sd, cfg, {"state_images": inp["state_images"][:n], "modality": "lang"}, inp["x_T"][:n], inp["goal"][:n], sg, name)
                g = gcs.SAMPLERS[name](model, {"state_images": inp["state_images"][:n].to(dev), "modality": "lang"}, inp["x_T"][:n].to(dev),
                                       inp["goal"][:n].to(dev), sg.to(dev)).cpu()
                res[f"{tag}_{name}"] = float((g - w).abs().max())
        except Exception as e:  # noqa: BLE001
            res[tag + "_error"] = repr(e)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
