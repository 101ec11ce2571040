"""The four decoder GEMMs at the per-branch M the B=256 sampling graph runs (640 rows), two launches each, for
   ncu --set full --clock-control none --import-source on -k regex:tc_gemm -o gpurun_out/r02_gemm python tools/ncu_gemm.py
The capture feeds profiles/r02_gemm_traffic.json (bench.py's roofline.traffic)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import helpers

B = 256
model = helpers.build_product(helpers.mdtv_inner_cfg(), 0, "trained")
state = {"state_images": torch.randn(B, 2, 384, device="cuda"), "modality": "lang"}
with torch.no_grad():     # the inference engine (with grad enabled the call would take the training path and create no engine)
    model(state, torch.randn(B, 10, 7, device="cuda"), torch.randn(B, 1, 512, device="cuda"), torch.ones(B, device="cuda"))
eng = list(model.inner_model._engines.values())[0]
M, d = 640, 384
for m, n, k, epi in ((M, 3 * d, d, 0), (M, d, d, 5), (M, 4 * d, d, 1), (M, d, 4 * d, 5)):
    eng.gemm_time_us(m, n, k, epi, 1)
torch.cuda.synchronize()
