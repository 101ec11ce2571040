"""Training-step timing (bench.train_measure: eager + graph-captured) or, with `ncu`, a few eager steps for a launch list:
   python tools/train_once.py            -> JSON record
   ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv python tools/train_once.py eager 3"""
import json, math, os, sys, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

args = types.SimpleNamespace(layers=4, precision="bf16x3")
if len(sys.argv) > 1 and sys.argv[1] == "eager":
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
    inp = synthetic_inputs(B, seed=31)
    torch.manual_seed(0)
    sig = U.rand_log_logistic((B,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0, device="cpu").cuda()
    batch = {k: inp[k].cuda() for k in ("state_images", "goal", "actions", "noise")}
    state = {"state_images": batch["state_images"], "modality": "lang"}
    for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
        if it == (int(sys.argv[2]) if len(sys.argv) > 2 else 3) - 1:
            torch.cuda.synchronize(); print("LAST STEP BEGINS", flush=True)
        opt.zero_grad(set_to_none=True)
        loss, _ = model.loss(state, batch["actions"], batch["goal"], batch["noise"], sig)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    print("loss", float(loss))
else:
    print(json.dumps(bench.train_measure(args, torch.device("cuda"), 512, 20, 5)))
