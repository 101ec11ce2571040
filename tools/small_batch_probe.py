"""10-step DDIM call latency (ms, median of 30) for small batches; run once per MDTB200_FUSED setting:
   for f in 0 1; do MDTB200_FUSED=$f python tools/small_batch_probe.py; done"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mdt_policy_b200 import GCDenoiser, DenoiseAgent
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs

dev = torch.device("cuda")
model = GCDenoiser(bench.inner_cfg(4, 4, "bf16x3", 64), sigma_data=0.5)
model.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], 12, "trained"))
model = model.to(dev).eval()
agent = DenoiseAgent(model, device=dev, num_sampling_steps=10, sampler_type="ddim", noise_scheduler="exponential", sigma_min=0.001, sigma_max=80.0)
res = {}
ref = {}
for b in (1, 4, 12, 24, 36, 64):
    inp = synthetic_inputs(b, seed=24)
    state = {"state_images": inp["state_images"].to(dev), "modality": "lang"}
    goal, xT = inp["goal"].to(dev), inp["x_T"].to(dev)
    for _ in range(5):
        out = agent.denoise_actions(None, state, goal, inference=True, x_T=xT)
    torch.cuda.synchronize()
    ms = []
    for _ in range(30):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); agent.denoise_actions(None, state, goal, inference=True, x_T=xT); e.record(); e.synchronize()
        ms.append(s.elapsed_time(e))
    ms.sort()
    res[b] = round(ms[15], 3)
    ref[b] = float(out.abs().sum())
print("FUSED=%s" % os.environ.get("MDTB200_FUSED", "0"), res, "checksums", {k: round(v, 4) for k, v in ref.items()})
