"""Debug aid for the persistent fused decoder: runs the same sampling call through the per-kernel path (MDTB200_FUSED=0) and
the fused kernel and compares the final actions and the intermediate buffers the engine exposes (debug_copy)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers as H  # noqa: E402
from mdt_policy_b200 import gc_sampling as gcs  # noqa: E402
from mdt_policy_b200.synthetic import synthetic_inputs  # noqa: E402


def run(fused, B, n_dec, n_steps, sampler="ddim", bufs=("qkv", "xh", "x")):
    # reference = per-kernel path without the algebraic cross-attention; candidate chosen by DEBUG_MODE (fused | cross)
    mode = os.environ.get("DEBUG_MODE", "cross")
    os.environ["MDTB200_FUSED"] = "1" if (fused and mode == "fused") else "0"
    os.environ["MDTB200_CROSS_FUSED"] = "1" if (fused and mode == "cross") else "0"
    model = H.build_product(H.mdtv_inner_cfg(2, n_dec, precision="bf16x3", ), 3, "trained")
    model.inner_model.max_batch = max(B, 16)
    inp = {k: v.cuda() for k, v in synthetic_inputs(B, seed=4).items()}
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sig = gcs.get_sigmas_exponential(n_steps, 0.001, 80.0, "cuda")
    fn = {"ddim": gcs.sample_ddim, "euler": gcs.sample_euler, "heun": gcs.sample_heun, "dpmpp_2m": gcs.sample_dpmpp_2m}[sampler]
    torch.cuda.synchronize()
    t0 = time.time()
    out = fn(model, state, inp["x_T"], inp["goal"], sig, disable=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    eng = list(model.inner_model._engines.values())[0]
    d = 384
    sizes = {"qkv": B * 10 * 3 * d, "q": B * 10 * d, "xh": B * 10 * d, "x": B * 10 * 7}
    dump = {n: eng.debug_buffer(n, sizes[n]).clone() for n in bufs}
    return out.cpu(), {k: v.cpu() for k, v in dump.items()}, dt


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    n_dec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    n_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sampler = sys.argv[4] if len(sys.argv) > 4 else "ddim"
    ref, rb, _ = run(False, B, n_dec, n_steps, sampler)
    got, gb, dt = run(True, B, n_dec, n_steps, sampler)
    print(f"B={B} n_dec={n_dec} n_steps={n_steps} {sampler}: final |fused - unfused| = {float((ref - got).abs().max()):.3e} (|ref| max {float(ref.abs().max()):.3f}), first fused call {dt*1e3:.1f} ms")
    for k in rb:
        dlt = (rb[k] - gb[k]).abs()
        print(f"  buffer {k:4s}: max diff {float(dlt.max()):.3e} at {int(dlt.argmax())} of {dlt.numel()}  (ref max {float(rb[k].abs().max()):.3f})")
