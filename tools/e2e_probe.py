"""Where does the e2e (host-buffer) call spend its extra ~0.4 ms?  Times pieces with CUDA events + wall clock."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import inner_cfg
from mdt_policy_b200 import GCDenoiser, DenoiseAgent
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs

dev = torch.device("cuda", 0)
B = 256
model = GCDenoiser(inner_cfg(4, 4, "bf16x3", B), sigma_data=0.5)
model.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], 12, "trained"))
model = model.to(dev).eval()
agent = DenoiseAgent(model, device=dev, num_sampling_steps=10, sampler_type="ddim", sigma_min=0.001, sigma_max=80.0)
inp = synthetic_inputs(B, seed=22)
host = {k: inp[k].pin_memory() for k in ("state_images", "goal", "x_T")}
d_state = {"state_images": host["state_images"].to(dev), "modality": "lang"}
d_goal, d_xT = host["goal"].to(dev), host["x_T"].to(dev)
out_host = torch.empty(B, 10, 7, pin_memory=True)

def ev_time(fn, k=50, flush=None):
    ts, ws = [], []
    for _ in range(k):
        if flush is not None: flush.zero_()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter(); s.record(); fn(); e.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
        ts.append(s.elapsed_time(e)); ws.append((w1 - w0) * 1e3)
    ts.sort(); ws.sort()
    return ts[len(ts) // 2], ws[len(ws) // 2]

flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
dev_call = lambda: agent.denoise_actions(None, d_state, d_goal, inference=True, x_T=d_xT)
host_call = lambda: agent.denoise_actions_host(host["state_images"], host["goal"], host["x_T"], "lang", out_host)
copies = lambda: (host["state_images"].to(dev, non_blocking=True), host["goal"].to(dev, non_blocking=True), host["x_T"].to(dev, non_blocking=True),
                  out_host.copy_(d_xT, non_blocking=True))
for _ in range(5): dev_call(); host_call()
for name, fn in (("device call", dev_call), ("host call", host_call), ("H2D+D2H copies only", copies)):
    for fl in (None, flush):
        t, w = ev_time(fn, flush=fl)
        print(f"{name:22s} flush={'yes' if fl is not None else 'no '}: events {t:.3f} ms, wall {w:.3f} ms")

# raw C-ABI host call (no Python agent logic around it)
import ctypes as C
inner = model.inner_model
eng = list(inner._engines.values())[0]
sig_host = agent._schedule_cache[("host", 10, "exponential", 0.001, 80.0)]
goal2d = host["goal"].reshape(B, -1)
def raw_host():
    out_host.copy_(host["x_T"])
    eng.lib.mdtb200_sample_host(eng.handle, 0, C.c_void_p(sig_host.data_ptr()), 10, C.c_void_p(goal2d.data_ptr()),
                                C.c_void_p(host["state_images"].data_ptr()), 1, B, C.c_void_p(out_host.data_ptr()), eng.stream)
x_dev = d_xT.clone()
sig_dev = sig_host.to(dev)
def raw_dev():
    eng.lib.mdtb200_sample(eng.handle, 0, C.c_void_p(sig_dev.data_ptr()), 10, C.c_void_p(d_goal.data_ptr()),
                           C.c_void_p(d_state["state_images"].data_ptr()), 1, B, C.c_void_p(x_dev.data_ptr()), eng.stream)
for name, fn in (("raw C host call", raw_host), ("raw C device call", raw_dev)):
    t, w = ev_time(fn, flush=flush)
    print(f"{name:22s} flush=yes: events {t:.3f} ms, wall {w:.3f} ms")
# cost of the graph launch call itself on the host
torch.cuda.synchronize()
w0 = time.perf_counter()
for _ in range(20): raw_dev()
w1 = time.perf_counter(); torch.cuda.synchronize(); w2 = time.perf_counter()
print(f"host time to ENQUEUE one device call: {(w1 - w0) / 20 * 1e3:.3f} ms; total per call when pipelined: {(w2 - w0) / 20 * 1e3:.3f} ms")
