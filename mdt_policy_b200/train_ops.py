"""Thin Python wrappers over the fused training primitives of libmdtb200.so (include/mdtb200.h, "training primitives, fused").

Operands: every tensor that enters a tensor-core GEMM is held as ONE row-major hi|lo bf16 split (``[rows, 2 * cols]`` bfloat16:
columns ``[0, cols)`` = hi, ``[cols, 2 cols)`` = lo, hi + lo = the fp32 value to ~2^-17).  The same array serves the forward
(K-major), the input gradient (weights read MN-major) and the weight gradient (both operands MN-major), so nothing is transposed.
All functions launch on torch's current stream and allocate their outputs with torch (plumbing); no CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ACT_GELU, ACT_MISH, ACT_SILU = 1, 2, 3
EPI_NONE, EPI_GELU16, EPI_GELUBWD16 = 0, 6, 7


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    # raw handle of torch's current stream on the tensor's device (torch.cuda.current_stream() builds a Stream object: ~7 us per call,
    # 300+ calls per training step)
    return _lib.current_stream_ptr(t.device.index)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {_lib.load().mdtb200_last_error(None).decode()}")


def group_sum(src2d, G, T, out=None):
    """out[g, c] = sum_t src[g*T + t, c]"""
    Cn = src2d.shape[-1]
    if out is None:
        out = torch.empty(G, Cn, dtype=torch.float32, device=src2d.device)
    _chk(_lib.load().mdtb200_op_group_sum(_p(src2d), _p(out), G, T, Cn, 0, _stream(src2d)), "op_group_sum")
    return out


def split(x2d, h=None, act=0, want_colsum=False):
    """x (M, K) fp32 -> (M, 2K) bf16 hi|lo of x * act'(h); optionally also the column sums (bias gradient)."""
    M, K = x2d.shape
    lib = _lib.load()
    out16 = torch.empty(M, 2 * K, dtype=torch.bfloat16, device=x2d.device)
    slabs = (M + lib.mdtb200_op_split_rows_per_slab() - 1) // lib.mdtb200_op_split_rows_per_slab()
    partial = torch.empty(slabs, K, dtype=torch.float32, device=x2d.device) if want_colsum else None
    _chk(lib.mdtb200_op_split(_p(x2d), _p(h), act, _p(out16), _p(partial), M, K, _stream(x2d)), "op_split")
    if want_colsum:
        return out16, group_sum(partial, 1, slabs).view(K)
    return out16


class SplitKWorkspace:
    """fp32 partial-tile workspace + per-tile arrival counters for the deterministic split-K GEMMs, one per device and stream role.
    Split-K launches that share a workspace must be ordered on ONE stream (they are: slot 0 = the caller's stream, slot 1 = the
    weight-gradient side stream); training two models concurrently from two host threads on the same device is not supported."""
    _inst = {}

    def __init__(self, device):
        self.ws = torch.empty(0, dtype=torch.float32, device=device)
        self.cnt = torch.zeros(4096, dtype=torch.int32, device=device)

    @classmethod
    def get(cls, device, floats, slot=0):
        """slot 0: GEMMs on the caller's stream; slot 1: the weight-gradient side stream (they run concurrently)"""
        key = (slot, device.index if device.index is not None else torch.cuda.current_device())
        inst = cls._inst.get(key)
        if inst is None:
            inst = cls._inst[key] = cls(device)
        if inst.ws.numel() < floats:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("split-K workspace must be sized before CUDA-graph capture (run one eager step first)")
            if slot == 1 and inst.ws.numel() > 0:
                inst.ws.record_stream(_side(device))      # kernels on the side stream may still be using the old buffer
            inst.ws = torch.empty(int(floats * 1.25), dtype=torch.float32, device=device)
        return inst


def _pick_splits(rows, cols, reduce):
    """split-K factor for an output of rows x cols (tiles of 128 x 128|64) reduced over `reduce`: fill ~1.4 waves of the 148 SMs,
    keep >= 4 k-blocks of 64 per slice"""
    tiles = ((rows + 127) // 128) * (cols // (128 if cols % 128 == 0 else 64))
    nkb = (reduce + 63) // 64
    return int(max(1, min(8, round(200 / max(tiles, 1)), nkb // 4)))


def gemm16(mode, a16, b16, M, N, K, bias=None, epi=EPI_NONE, want16=False, splits=None):
    """mode 0: (M,N) = x16 . w16^T + bias;  mode 1: (M,K) = dy16 . w16;  mode 2: (N,K) = dy16^T . x16 (split-K over M)."""
    lib = _lib.load()
    dev = a16.device
    shape = (M, N) if mode == 0 else (M, K) if mode == 1 else (N, K)
    out = torch.empty(shape, dtype=torch.float32, device=dev)
    out16 = torch.empty(M, 2 * N, dtype=torch.bfloat16, device=dev) if (want16 or epi == EPI_GELU16) else None
    ws = cnt = None
    s = 1
    if mode == 2 or (mode == 1 and N >= 2048):       # long reductions with few output tiles: wgrad (over M), dgrad of the stacked groups (over N)
        s = (_pick_splits(N, K, M) if mode == 2 else _pick_splits(M, K, N)) if splits is None else splits
        if s > 1:
            w = SplitKWorkspace.get(dev, lib.mdtb200_op_gemm16_ws(N, K, s) if mode == 2 else lib.mdtb200_op_gemm16_ws(M, K, s))
            ws, cnt = w.ws, w.cnt
    _chk(lib.mdtb200_op_gemm16(mode, _p(a16), _p(b16), _p(bias), _p(out), _p(out16), M, N, K, epi, s, _p(ws), _p(cnt), _stream(a16)), "op_gemm16")
    return (out, out16) if out16 is not None else out


# Weight-gradient GEMMs have no consumer inside the backward pass, so they run on a side stream next to the (latency-bound) chain of
# input-gradient kernels of the same residual branch and are joined before the branch's backward returns (autograd / DDP hooks see
# finished gradients).  This pays off when the GPU is the bottleneck, i.e. while the step is being captured into a CUDA graph
# (GraphedTrainStep: 4.97 -> 4.55 ms); an eager step is host-bound and the extra event traffic costs ~1 ms, so by default the side
# stream is used only under capture.  MDTB200_TRAIN_WGRAD_STREAM=0 / 1 forces it off / on everywhere.
import os as _os
_WGRAD_MODE = _os.environ.get("MDTB200_TRAIN_WGRAD_STREAM", "capture")
_side_streams = {}


def _use_side():
    if _WGRAD_MODE == "0":
        return False
    return _WGRAD_MODE == "1" or torch.cuda.is_current_stream_capturing()


def _side(dev):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=dev)
    return st


def wgrad_begin(dy16, x16, M, N, K):
    """dW (N, K) = dy16^T . x16, launched on the side stream; call wgrad_join(device) before the result leaves the backward."""
    if not _use_side():
        return gemm16(2, dy16, x16, M, N, K)
    dev = dy16.device
    lib = _lib.load()
    out = torch.empty(N, K, dtype=torch.float32, device=dev)
    s = _pick_splits(N, K, M)
    ws = cnt = None
    if s > 1:
        w = SplitKWorkspace.get(dev, lib.mdtb200_op_gemm16_ws(N, K, s), slot=1)
        ws, cnt = w.ws, w.cnt
    side = _side(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    _chk(lib.mdtb200_op_gemm16(2, _p(dy16), _p(x16), None, _p(out), None, M, N, K, EPI_NONE, s, _p(ws), _p(cnt), C.c_void_p(side.cuda_stream)), "op_gemm16 (wgrad)")
    return out


def wgrad_join(dev):
    if _use_side():
        torch.cuda.current_stream(dev).wait_stream(_side(dev))


def presize_side_workspace(dev):
    """give the side-stream split-K workspace the size the eager warm-up steps needed on the main stream (call before graph capture)"""
    key0 = (0, dev.index if dev.index is not None else torch.cuda.current_device())
    inst = SplitKWorkspace._inst.get(key0)
    if inst is not None and inst.ws.numel() > 0:
        SplitKWorkspace.get(dev, inst.ws.numel(), slot=1)


def dgrad_gelu_bwd16(dy16, w16, h, M, N, K, want_colsum=False):
    """dh16 (M, 2K) = split((dy16 . w16) * GELU'(h)) in ONE GEMM (activation backward in the epilogue) [+ its column sums = bias gradient]"""
    lib = _lib.load()
    out16 = torch.empty(M, 2 * K, dtype=torch.bfloat16, device=dy16.device)
    mt = (M + 127) // 128
    part = torch.empty(mt, K, dtype=torch.float32, device=dy16.device) if want_colsum else None
    _chk(lib.mdtb200_op_gemm16(1, _p(dy16), _p(w16), _p(h), _p(part), _p(out16), M, N, K, EPI_GELUBWD16, 1, None, None, _stream(dy16)), "op_gemm16 (gelu bwd)")
    if want_colsum:
        return out16, group_sum(part, 1, mt).view(K)
    return out16


def ln_fwd16(x, w, b, shift, scale, mod_stride, T, want32=False):
    """x (M, d) -> split-bf16 LayerNorm(+modulate) output (M, 2d) [and the fp32 one]."""
    M, d = x.shape
    y16 = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=x.device)
    y = torch.empty(M, d, dtype=torch.float32, device=x.device) if want32 else None
    _chk(_lib.load().mdtb200_op_ln_fwd16(_p(x), _p(w), _p(b), _p(shift), _p(scale), mod_stride, T, M, d, _p(y), _p(y16), _stream(x)), "op_ln_fwd16")
    return (y16, y) if want32 else y16


def ln_bwd2(x, dy, w, b, scale, mod_stride, dres, dshift, dscale, dmod_stride, T):
    """-> dx (= dres + LN backward), d ln.weight, d ln.bias; dshift / dscale rows are written in place when given."""
    M, d = x.shape
    nb = _lib.load().mdtb200_op_ln_bwd2_partials(M, T)
    dx = torch.empty_like(x)
    partial = torch.empty(nb, 2 * d, dtype=torch.float32, device=x.device)
    _chk(_lib.load().mdtb200_op_ln_bwd2(_p(x), _p(dy), _p(w), _p(b), _p(scale), mod_stride, _p(dres), _p(dx), _p(dshift), _p(dscale), dmod_stride,
                                        _p(partial), M, d, T, _stream(x)), "op_ln_bwd2")
    wb = group_sum(partial, 1, nb).view(2 * d)
    return dx, wb[:d], (wb[d:] if b is not None else None)


def ln_fwd(x, w, b, T=1):
    """plain fp32 LayerNorm output (final norms)"""
    M, d = x.shape
    y = torch.empty_like(x)
    _chk(_lib.load().mdtb200_op_ln_fwd16(_p(x), _p(w), _p(b), None, None, 0, T, M, d, _p(y), None, _stream(x)), "op_ln_fwd16")
    return y


def attn_fwd16(q, ldq, k, v, ldkv, B, H, hd, Tq, Tk, causal, p_drop, seed):
    y16 = torch.empty(B * Tq, 2 * H * hd, dtype=torch.bfloat16, device=q.device)
    _chk(_lib.load().mdtb200_op_attn_fwd16(_p(q), ldq, _p(k), _p(v), ldkv, _p(y16), B, H, hd, Tq, Tk, int(causal), float(p_drop), seed, _stream(q)),
         "op_attn_fwd16")
    return y16


def attn_bwd(q, ldq, k, v, ldkv, dy, dq, lddq, dk, dv, lddkv, B, H, hd, Tq, Tk, causal, p_drop, seed):
    _chk(_lib.load().mdtb200_op_attn_bwd(_p(q), ldq, _p(k), _p(v), ldkv, _p(dy), H * hd, _p(dq), lddq, _p(dk), _p(dv), lddkv, B, H, hd, Tq, Tk,
                                         int(causal), float(p_drop), seed, _stream(q)), "op_attn_bwd")


ATTN_BWD16_SHAPES = {(48, 10, 10, True), (48, 10, 4, True), (48, 4, 4, False), (64, 10, 10, True), (64, 10, 3, True), (64, 3, 3, False)}


def attn_bwd16_self(qkv, dy, B, H, hd, T, causal, p_drop, seed, want_bias):
    """self-attention backward -> dqkv16 (M, 2*3D) operand of the fused q|k|v projection [+ (3D,) bias gradient]"""
    D = H * hd
    M = B * T
    d16 = torch.empty(M, 6 * D, dtype=torch.bfloat16, device=qkv.device)
    part = torch.empty(B, 3 * D, dtype=torch.float32, device=qkv.device) if want_bias else None
    _chk(_lib.load().mdtb200_op_attn_bwd16(_p(qkv), 3 * D, _p(qkv[:, D:]), _p(qkv[:, 2 * D:]), 3 * D, _p(dy), D, None, None, 0, _p(d16), 3 * D, 0,
                                           _p(d16), 3 * D, D, 2 * D, _p(part), 1, B, H, hd, T, T, int(causal), float(p_drop), seed, _stream(qkv)),
         "op_attn_bwd16")
    return d16, (group_sum(part, 1, B).view(3 * D) if want_bias else None)


def attn_bwd16_cross(q, kv, dy, B, H, hd, Tq, Tk, causal, p_drop, seed, want_bias):
    """cross-attention backward -> dq16 (M, 2D) operand of the query projection, dkv (B, Tk, 2D) fp32 [+ (D,) query bias gradient]"""
    D = H * hd
    M = B * Tq
    d16 = torch.empty(M, 2 * D, dtype=torch.bfloat16, device=q.device)
    dkv = torch.empty(B, Tk, 2 * D, dtype=torch.float32, device=q.device)
    part = torch.empty(B, D, dtype=torch.float32, device=q.device) if want_bias else None
    _chk(_lib.load().mdtb200_op_attn_bwd16(_p(q), D, _p(kv), _p(kv[..., D:]), kv.stride(1), _p(dy), D, _p(dkv), _p(dkv[..., D:]), 2 * D, _p(d16), D, 0,
                                           None, 0, 0, 0, _p(part), 0, B, H, hd, Tq, Tk, int(causal), float(p_drop), seed, _stream(q)),
         "op_attn_bwd16")
    return d16, dkv, (group_sum(part, 1, B).view(D) if want_bias else None)


def res_drop_fwd(x, f, gate, gate_stride, T, p, seed):
    M, d = x.shape
    out = torch.empty_like(x)
    _chk(_lib.load().mdtb200_op_res_drop_fwd(_p(x), _p(f), _p(gate), gate_stride, _p(out), M, d, T, float(p), seed, _stream(x)), "op_res_drop_fwd")
    return out


def res_drop_bwd(dout, f, gate, gate_stride, dgate, dgate_stride, T, p, seed, want_bias=False):
    M, d = dout.shape
    G = (M + T - 1) // T
    df16 = torch.empty(M, 2 * d, dtype=torch.bfloat16, device=dout.device)
    bpartial = torch.empty(G, d, dtype=torch.float32, device=dout.device) if want_bias else None
    _chk(_lib.load().mdtb200_op_res_drop_bwd(_p(dout), _p(f), _p(gate), gate_stride, _p(df16), _p(dgate), dgate_stride, _p(bpartial), M, d, T,
                                             float(p), seed, _stream(dout)), "op_res_drop_bwd")
    if want_bias:
        return df16, group_sum(bpartial, 1, G).view(d)
    return df16


def narrow_fwd(x2d, W, bias):
    M, K = x2d.shape
    J = W.shape[0]
    y = torch.empty(M, J, dtype=torch.float32, device=x2d.device)
    _chk(_lib.load().mdtb200_op_narrow_fwd(_p(x2d), _p(W), _p(bias), _p(y), M, K, J, _stream(x2d)), "op_narrow_fwd")
    return y


def narrow_wgrad(wide, thin, wide_major):
    """sum_m wide[m, n] thin[m, j] -> (N, J) if wide_major else (J, N)"""
    M, N = wide.shape
    J = thin.shape[1]
    slabs = (M + 63) // 64
    partial = torch.empty(slabs, N * J, dtype=torch.float32, device=wide.device)
    _chk(_lib.load().mdtb200_op_narrow_wgrad(_p(wide), _p(thin), _p(partial), M, N, J, int(wide_major), _stream(wide)), "op_narrow_wgrad")
    out = group_sum(partial, 1, slabs)
    return out.view(N, J) if wide_major else out.view(J, N)


class WeightBank:
    """All tensor-core weight operands of a network, refreshed by ONE kernel launch per step.

    ``groups``: list of (name, [weights (N_i, K)], [biases or None]) -- the weights of a group are stacked row-wise into one
    operand ``w16[name]`` of shape (sum N_i, 2K) (q/k/v projections, the cross-attention K/V of all layers, the AdaLN modulation of
    all layers), their biases into one fp32 vector ``bias[name]`` (None when the group has no bias).
    """

    def __init__(self, groups, device):
        import struct
        self.device = device
        self.w16, self.bias, self.params = {}, {}, []
        n16 = sum(sum(w.numel() for w in ws) * 2 for _, ws, _ in groups)
        nb = sum(sum(b.numel() for b in bs) for _, _, bs in groups if bs[0] is not None)
        self.arena16 = torch.empty(n16 + 64, dtype=torch.bfloat16, device=device)
        self.arena32 = torch.empty(nb + 64, dtype=torch.float32, device=device)
        recs, blocks, o16, o32 = [], [], 0, 0
        for name, ws, bs in groups:
            K = ws[0].shape[1]
            rows = sum(w.shape[0] for w in ws)
            self.w16[name] = self.arena16[o16:o16 + rows * 2 * K].view(rows, 2 * K)
            r = 0
            for w in ws:
                if w.shape[1] != K or not w.is_contiguous() or w.dtype != torch.float32 or K % 4:
                    raise RuntimeError(f"weight group '{name}': members must be contiguous fp32 (N_i, {K}) matrices")
                dst = self.arena16.data_ptr() + (o16 + r * 2 * K) * 2
                blocks += [(len(recs), c) for c in range((w.numel() + 4095) // 4096)]
                recs.append(struct.pack("<QQqii", w.data_ptr(), dst, w.numel(), K, 0))
                self.params.append(w)
                r += w.shape[0]
            o16 += rows * 2 * K
            o16 = (o16 + 7) // 8 * 8
            if bs[0] is not None:
                n = sum(b.numel() for b in bs)
                self.bias[name] = self.arena32[o32:o32 + n]
                r = 0
                for b in bs:
                    if b.numel() % 4 or not b.is_contiguous():
                        raise RuntimeError(f"weight group '{name}': biases must be contiguous with a multiple of 4 elements")
                    blocks += [(len(recs), c) for c in range((b.numel() + 4095) // 4096)]
                    recs.append(struct.pack("<QQqii", b.data_ptr(), self.arena32.data_ptr() + (o32 + r) * 4, b.numel(), 0, 0))
                    self.params.append(b)
                    r += b.numel()
                o32 = (o32 + n + 3) // 4 * 4
            else:
                self.bias[name] = None
        self.ptrs = tuple(p.data_ptr() for p in self.params)
        self.table = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(device)
        self.blocks = torch.tensor(blocks, dtype=torch.int32).to(device)
        self.n_blocks = len(blocks)

    def valid(self):
        return self.ptrs == tuple(p.data_ptr() for p in self.params)

    def refresh(self):
        _chk(_lib.load().mdtb200_op_split_multi(_p(self.table), _p(self.blocks), self.n_blocks, _stream(self.arena16)), "op_split_multi")
