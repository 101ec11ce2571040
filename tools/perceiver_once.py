import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdt_policy_b200.perceiver import PerceiverResampler
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
m = PerceiverResampler(dim=384, depth=6, dim_head=64, heads=8, num_latents=3, num_time_embeds=1, max_batch=B).cuda()
x = torch.randn(B, 1, 392, 384, device="cuda")
for _ in range(3):
    m(x)
torch.cuda.synchronize()
print("ok")
