"""cProfile of the eager training step (host side): python tools/profile_eager_step.py"""
import cProfile, math, os, pstats, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mdt_policy_b200 import GCDenoiser, utils as U
from mdt_policy_b200.optim import FusedAdamWEMA
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs
B = 512
cfgd = bench.inner_cfg(4, 4, "fp32", B)
cfgd.update(dict(attn_pdrop=0.3, resid_pdrop=0.1, mlp_pdrop=0.05))
model = GCDenoiser(cfgd, sigma_data=0.5)
model.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], 12, "trained"))
model = model.cuda().train()
opt = FusedAdamWEMA(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.999)
inp = {k: v.cuda() for k, v in synthetic_inputs(B, seed=31).items()}
sig = U.rand_log_logistic((B,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0, device="cpu").cuda()
state = {"state_images": inp["state_images"], "modality": "lang"}
def step():
    opt.zero_grad(set_to_none=True)
    loss, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sig)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
