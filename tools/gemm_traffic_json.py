"""Turns the `ncu --set full` capture of tools/ncu_gemm.py into profiles/r02_gemm_traffic.json (bench.py's roofline.traffic):
   python tools/gemm_traffic_json.py gpurun_out/r02_gemm.ncu-rep
Reads the report with `ncu -i ... --page raw --csv` (works without a GPU).  Launch order = tools/ncu_gemm.py: the four decoder GEMMs at
the per-chain M = 640, two launches each, matched by output width (grid.x * tile width) and, for the two d-wide
GEMMs, by duration (K = 4d is the longer one)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ns": 1e-3, "ms": 1e3, "%": 1.0}


def val(r, name):
    i = col[name]
    return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)


M, d = 640, 384
shapes = [("qkv (N=3d,K=d)", 3 * d, d, "bias"), ("attn c_proj + gate + res (N=d,K=d)", d, d, "res"),
          ("mlp c_fc + GELU (N=4d,K=d)", 4 * d, d, "split"), ("mlp c_proj + gate + res (N=d,K=4d)", d, 4 * d, "res")]
gemm = [r for r in data if "tc_gemm_kernel" in r[col["Kernel Name"]]]
gemm = gemm[-8:]          # the eight launches of tools/ncu_gemm.py (anything before is the warm-up forward)


def out_cols(r):          # N of a launch = grid.x * tile width (template argument BN)
    gx = int(r[col["Grid Size"]].strip("()").split(",")[0])
    return gx * int(r[col["Kernel Name"]].split("<")[1].split(",")[0])


by_n = {}
for r in gemm:
    by_n.setdefault(out_cols(r), []).append(r)
dwide = sorted(by_n[d], key=lambda r: val(r, "gpu__time_duration.sum"))      # N = d: the K = d launches are the shorter half
pick = {0: by_n[3 * d], 1: dwide[:len(dwide) // 2], 2: by_n[4 * d], 3: dwide[len(dwide) // 2:]}
out = {"shapes": {}}
tot_bytes = tot_n = 0
for si, (name, N, K, kind) in enumerate(shapes):
    rs = pick[si]
    grid = rs[0][col["Grid Size"]]
    gx, gy = [int(x) for x in grid.strip("()").split(",")[:2]]
    rd = sum(val(r, "dram__bytes_read.sum") for r in rs) / len(rs)
    wr = sum(val(r, "dram__bytes_write.sum") for r in rs) / len(rs)
    # algorithmic bytes: split-bf16 operands (4 B per element of A and W) + fp32 bias (+ fp32 residual read) ; output fp32 or split bf16 (4 B)
    alg_rd = 4 * (M * K + N * K) + 4 * N + (4 * M * N if kind == "res" else 0)
    alg_wr = 4 * M * N
    out["shapes"][name] = {
        "grid": gx * gy, "tile_n": N // gx,
        "ncu_us": [round(val(r, "gpu__time_duration.sum"), 3) for r in rs],
        "dram_read_bytes": rd, "dram_write_bytes": wr,
        "algorithmic_read_bytes": alg_rd, "algorithmic_write_bytes": alg_wr,
        "tensor_pipe_pct": [val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for r in rs],
        "l2_to_sm_bytes": sum(val(r, "l1tex__m_xbar2l1tex_read_bytes.sum") for r in rs) / len(rs),
    }
    tot_bytes += rd + wr
    tot_n += 1
out = {"dram_bytes_per_launch": tot_bytes / tot_n,
       "source": f"profiles/r02_gemm_traffic.json <- ncu --set full --clock-control none capture {os.path.relpath(rep, ROOT)} (tools/ncu_gemm.py: the four decoder "
                 "GEMMs at the per-branch M=640 of B=256 / 4 chains, two launches each; tools/gemm_traffic_json.py); dram__bytes_read.sum + dram__bytes_write.sum, "
                 "launch-weighted mean",
       "note": "writes are 0 because the outputs stay in the 126 MB L2 under ncu's serialised replay; reads equal the algorithmic operand bytes within 20 % "
               "(cold L2: operands + bias/gate vectors + descriptors)",
       **out}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_gemm_traffic.json"), "w"), indent=1)
print(json.dumps({k: (v["grid"], v["tile_n"], v["ncu_us"], round(v["l2_to_sm_bytes"] / 1e6, 2)) for k, v in out["shapes"].items()}, indent=1))
print("dram_bytes_per_launch", out["dram_bytes_per_launch"])
