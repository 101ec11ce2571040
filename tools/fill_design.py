"""Regenerates the round-2 results table of DESIGN.md (between the R2TABLE markers) from a bench.py JSON line:
   python tools/fill_design.py profiles/r02_final_bench.json"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(sys.argv[1]))
cpu = d.get("cpu_baseline", {})
dk, rf, lat, tr, tb = d["dominant_kernel"], d["roofline"], d["latency"], d["train"], d["gpu_torch_baseline"]
sh = {k.split(" (")[0]: v["us_per_launch"] for k, v in dk["shapes"].items()}
calls = lat["calls"]
rows = [
    ("`value` (device-resident inputs, CUDA events, L2 flushed between timed calls)",
     f"**{d['value']:.0f} denoise-steps/s** = {d['ms_per_step']:.3f} ms per 10-step call (median {lat['device']['median']:.3f}, p95 {lat['device']['p95']:.3f} over {calls} calls)"),
    ("`e2e` (DenoiseAgent on pinned HOST tensors through `mdtb200_sample_host`, 1.78 MB H2D + 72 KB D2H inside the timed region)",
     f"**{d['e2e']['value']:.0f} steps/s** ({d['e2e']['ms_per_step']:.3f} ms; median {lat['e2e']['median']:.3f}, p95 {lat['e2e']['p95']:.3f})"),
    ("kernels per sampling call / per score evaluation", f"{d['gpu_launches'] // calls} / 30 (round 1: 1984 / 46)"),
    (f"top-level `roofline` = the WHOLE timed graph: ALG(B,N) = {rf['alg_gflop_per_launch']:.2f} GFLOP per launch / time / measured bf16 burst peak {rf['peak']:.1f} TFLOP/s",
     f"{rf['achieved']:.1f} TFLOP/s, frac **{rf['frac']:.4f}**"),
    ("`dominant_kernel` = `tc::tc_gemm_kernel`, timed live at the per-chain M = 640 the graph runs, tile width as `launch_tc_gemm` picks it, non-zero operands, launch-weighted over the 4 GEMMs of a layer",
     f"{dk['achieved']:.1f} TFLOP/s algorithmic, frac {dk['frac']:.4f} (" + ", ".join(f"{k} {v:.1f} us" for k, v in sh.items()) + f"); `traffic` {rf['traffic'] / 1e6:.2f} MB DRAM per launch = the algorithmic operand bytes (profiles/r02_gemm_traffic.json)"),
    ("`small_batch` (the weight-streaming / latency regime of SURVEY 8d; `hbm` roofline with algorithmic bytes = decoder weights + x + context per step)",
     " ; ".join(f"B={b}: {d['small_batch'][f'B{b}']['ms_per_call']:.2f} ms per 10-step call, {d['small_batch'][f'B{b}']['roofline']['achieved']:.0f} GB/s = frac {d['small_batch'][f'B{b}']['roofline']['frac']:.3f}" for b in (1, 32))
     + f" ; host (oracle port) B=1: {d['small_batch']['cpu_B1'].get('ms_per_call', float('nan')):.0f} ms") if "small_batch" in d and "B1" in d["small_batch"] else ("`small_batch`", "n/a"),
    ("`variant_6x6` (BASELINE-literal 6 enc + 6 dec)", f"{d['variant_6x6']['value']:.0f} steps/s ({d['variant_6x6']['ms_per_step']:.2f} ms)"),
    ("`gpu_torch_baseline`: the same algorithm as stock PyTorch on the same B200 (oracle port's ATen ops on cuda, encoder per evaluation like the reference)",
     f"eager fp32 {tb['eager_fp32']['value']:.0f} steps/s, CUDA-graphed fp32 {tb['graphed_fp32']['value']:.0f}, TF32 {tb['tf32']['value']:.0f} (action error {tb['tf32']['err']:.1e} vs the 1e-4 gate) -> this repo is {d['value'] / tb['graphed_fp32']['value']:.1f}x the graphed fp32 arm"),
    ("`train` (config 3: batch 512, loss fwd + bwd + fused AdamW/EMA, shipped dropout)",
     f"**{tr['graphed']['ms_per_step']:.2f} ms per step** as a replayed CUDA graph (`GraphedTrainStep`) = {tr['graphed']['value'] / 1e3:.0f} k action-tokens/s; {tr['eager']['ms_per_step']:.2f} ms eager (host-bound) (round 1: 13.9-15.7 ms)"),
    ("`perceiver` (8f-1: depth 6, 392 tokens -> 3 latents, B=256)", f"{d['perceiver']['ms_per_call']:.2f} ms per call, {d['perceiver']['launches_per_call']} launches"
     + (f"; the reference formulation as stock PyTorch fp32 on the same GPU: {d['perceiver']['gpu_torch_eager_fp32']['ms_per_call']:.1f} ms" if 'ms_per_call' in d['perceiver'].get('gpu_torch_eager_fp32', {}) else "")),
    ("CPU arm (`cpu_baseline`, oracle port, thread count calibrated)", f"{cpu.get('value', float('nan')):.1f} steps/s on {cpu.get('cores', '?')} threads -> e2e is {d['e2e']['value'] / cpu.get('value', float('nan')):.0f}x"),
    ("clocks during the timed region", f"{d['clocks']['sm_mhz']:.0f} MHz of {d['clocks']['sm_max_mhz']:.0f}, throttle reasons {d['clocks']['reasons']}"),
]
table = "| quantity | value |\n|---|---|\n" + "\n".join(f"| {a} | {b} |" for a, b in rows) + "\n"
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
s2 = re.sub(r"<!-- R2TABLE -->.*?<!-- /R2TABLE -->", "<!-- R2TABLE -->\n" + table + "<!-- /R2TABLE -->", s, flags=re.S)
assert s2 != s or "<!-- R2TABLE -->" in s
open(p, "w").write(s2)
print(table)
