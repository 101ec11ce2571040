"""Finds the first non-finite gradient / loss of the B=512 training loop (bench train workload)."""
import math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import inner_cfg
from mdt_policy_b200 import GCDenoiser, utils as U
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
drop = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cfgd = inner_cfg(4, 4, "fp32", B)
cfgd.update(dict(attn_pdrop=0.3 * drop, resid_pdrop=0.1 * drop, mlp_pdrop=0.05 * drop))
model = GCDenoiser(cfgd, sigma_data=0.5)
model.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], 12, "trained"))
model = model.cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=0.05)
inp = synthetic_inputs(B, seed=31)
torch.manual_seed(0)
sig = U.rand_log_logistic((B,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0, device="cpu").cuda()
batch = {k: inp[k].cuda() for k in ("state_images", "goal", "actions", "noise")}
state = {"state_images": batch["state_images"], "modality": "lang"}
print("sigma range", float(sig.min()), float(sig.max()))
nosync = len(sys.argv) > 4 and sys.argv[4] == "nosync"
if nosync:          # back-to-back steps without any host synchronisation (what a real training loop / bench.py does)
    losses = []
    for it in range(int(sys.argv[3])):
        opt.zero_grad(set_to_none=True)
        loss, out = model.loss(state, batch["actions"], batch["goal"], batch["noise"], sig)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
    torch.cuda.synchronize()
    print("nosync losses:", [round(float(l), 4) for l in losses])
    sys.exit(0)
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 8):
    opt.zero_grad(set_to_none=True)
    loss, out = model.loss(state, batch["actions"], batch["goal"], batch["noise"], sig)
    loss.backward()
    bad = [(n, float(p.grad.abs().max())) for n, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    gmax = max(float(p.grad.abs().max()) for p in model.parameters() if p.grad is not None and torch.isfinite(p.grad).all())
    if it % 5 == 0 or bad: print(f"it {it}: loss {float(loss):.5f} out finite {bool(torch.isfinite(out).all())} max finite grad {gmax:.3e} non-finite grads: {bad[:4]} ({len(bad)})")
    if bad or not math.isfinite(float(loss)):
        break
    opt.step()
    wbad = [n for n, p in model.named_parameters() if not torch.isfinite(p).all()]
    if wbad:
        print("non-finite weights after step:", wbad[:5]); break
