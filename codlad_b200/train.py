"""The train_latent step of the denoiser on the GPU (BASELINE.json configs[4]).

Reference: train_latent.py:184-261 -- `diffusion.training_losses(net_model, x1, t, dict(y, mask, batch))` (q_sample, one denoiser
forward in train mode, masked eps-MSE + VB term), `accelerator.backward` (DDP all-reduce of the 2 449 974 fp32 gradients),
`clip_grad_norm_(1.0)`, `AdamW.step`, `update_ema`.

Here the forward AND the backward of ProteinMPNN_diffusion_new (models/latent_model.py:175-268, models/protein_mpnn_utils.py:208-330,
447-523) are composed from the CUDA operators of include/codlad_b200_train.h (libcodlad_b200.so): fp32 GEMMs for every linear layer
and its two gradients, fused kernels for the gathers / GELU / masked neighbour sum / LayerNorm + adaLN modulation and their
gradients, deterministic reductions, one fused AdamW + EMA kernel over the flat parameter buffer.  W1 is applied in its factored
form (per-node products gathered per edge), so the 384 / 512-wide concatenations are never built -- forward or backward.

torch is used for what the task allows it: device memory (activations are torch CUDA tensors), integer index preparation of the
k-NN graph (neighbour node ids, the reverse CSR), the scalar loss of `training_losses` on the [B, L, 6] model output (its gradient
with respect to that output comes from torch.autograd on that tiny expression) and torch.distributed (NCCL) for the gradient
all-reduce.  No torch.nn layer and no torch matmul is on this path.

Dropout: the reference trains with p = 0.6 (latent_model.py:88); `dropout_p > 0` draws the masks with torch's generator and the
kernels apply them (mask / (1 - p)); gradient parity is tested at p = 0 (SURVEY.md 'hard parts').
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict

import torch

from . import _native as N
from . import batching, weights
from .engine import knn_topk, sinusoid_freqs

H = 128


def _p(t):
    return None if t is None else N.dptr(t)


class _Ops:
    """Typed wrappers over the cb2t_* entry points; tensors are fp32 CUDA, contiguous unless a leading dimension is given."""

    def __init__(self):
        self.lib = N.lib()
        self.transposed = {}           # data_ptr of a weight view -> its transposed copy (TF32 mode only)
        self.stream = N.stream_ptr()   # refreshed on entry of forward / backward (a pass makes ~1 500 calls)

    def gemm(self, A, B, Cm, M, Nn, K, lda, ldb, ldc, a_kc, b_kc, acc=False):
        N.check(self.lib.cb2t_gemm(A, B, Cm, int(M), int(Nn), int(K), int(lda), int(ldb), int(ldc), int(a_kc), int(b_kc), int(acc), self.stream),
                "gemm")

    # y[M, out] = x[M, in] @ W[out, in_total][:, c0:c0+in]^T   (W given as (tensor, col0, in))
    def linear(self, x, W, c0, kin, out=None, acc=False):
        M, nout, ldw = x.shape[0], W.shape[0], W.shape[1]
        y = out if out is not None else torch.empty(M, nout, device=x.device, dtype=torch.float32)
        self.gemm(x.data_ptr(), W.data_ptr() + 4 * c0, y.data_ptr(), M, nout, kin, x.stride(0), ldw, y.stride(0), 1, 1, acc)
        return y

    # dx[M, in] (+)= dy[M, out] @ W[:, c0:c0+in]
    def linear_dx(self, dy, W, c0, kin, out=None, acc=False):
        M, nout, ldw = dy.shape[0], W.shape[0], W.shape[1]
        dx = out if out is not None else torch.empty(M, kin, device=dy.device, dtype=torch.float32)
        Wt = self.transposed.get(W.data_ptr())
        if Wt is not None:             # TF32 mode: both operands K-contiguous (B = rows c0.. of W^T [in_total, out])
            self.gemm(dy.data_ptr(), Wt.data_ptr() + 4 * c0 * nout, dx.data_ptr(), M, kin, nout, dy.stride(0), nout, dx.stride(0), 1, 1, acc)
        else:
            self.gemm(dy.data_ptr(), W.data_ptr() + 4 * c0, dx.data_ptr(), M, kin, nout, dy.stride(0), ldw, dx.stride(0), 1, 0, acc)
        return dx

    # dW[:, c0:c0+in] += dy[M, out]^T @ x[M, in]
    def linear_dw(self, dy, x, dW, c0):
        M, nout, kin = dy.shape[0], dy.shape[1], x.shape[1]
        self.gemm(dy.data_ptr(), x.data_ptr(), dW.data_ptr() + 4 * c0, nout, kin, M, dy.stride(0), x.stride(0), dW.shape[1], 0, 0, True)

    def colsum(self, X, out, acc=True):
        N.check(self.lib.cb2t_colsum(_p(X), X.shape[0], X.shape[1], X.stride(0), _p(out), int(acc), self.stream), "colsum")

    def linear_bias_act(self, x, W, bias, c0, kin, want_act=True):
        """(Z, Y) = (x @ W[:, c0:c0+kin]^T + bias, GELU(Z) or None): one fused kernel in TF32 mode (cb2t_linear_bias_gelu_fwd)."""
        M, nout, ldw = x.shape[0], W.shape[0], W.shape[1]
        Z = torch.empty(M, nout, device=x.device, dtype=torch.float32)
        Y = torch.empty_like(Z) if want_act else None
        N.check(self.lib.cb2t_linear_bias_gelu_fwd(x.data_ptr(), W.data_ptr() + 4 * c0, _p(bias), _p(Z), _p(Y), M, nout, kin, x.stride(0), ldw, nout,
                                                   self.stream), "linear_bias_gelu_fwd")
        return Z, Y

    def bias_gelu(self, Z, bias, want_act=True):
        Y = torch.empty_like(Z) if want_act else None
        N.check(self.lib.cb2t_bias_gelu_fwd(_p(Z), _p(bias), Z.shape[0], Z.shape[1], _p(Y), self.stream), "bias_gelu_fwd")
        return Y

    def gelu_bwd(self, pre, dY, bias_grad=None):
        """dY <- dY * GELU'(pre) in place; bias_grad (+)= its column sums in the same pass (the bias gradient of the layer behind `pre`)."""
        if bias_grad is not None:
            N.check(self.lib.cb2t_gelu_bwd_colsum(_p(pre), _p(dY), pre.shape[0], pre.shape[1], _p(dY), _p(bias_grad), 1, self.stream), "gelu_bwd_colsum")
        else:
            N.check(self.lib.cb2t_gelu_bwd(_p(pre), _p(dY), pre.numel(), _p(dY), self.stream), "gelu_bwd")
        return dY

    def ew(self, mode, a, b=None, scale=1.0, out=None):
        out = out if out is not None else torch.empty_like(a)
        N.check(self.lib.cb2t_elementwise(mode, _p(a), _p(b), float(scale), a.numel(), _p(out), self.stream), "elementwise")
        return out

    def edge_combine_gelu(self, Z, Pa, Pc, bias, g):
        Y = torch.empty_like(Z)
        N.check(self.lib.cb2t_edge_combine_gelu_fwd(_p(Z), _p(Pa), _p(Pc), _p(bias), _p(g.nbr_node), g.K, Z.shape[0], _p(Y), self.stream), "edge_combine")
        return Y

    def edge_gather_bwd(self, dZ, g):
        dPa = torch.empty(g.Nn, H, device=dZ.device)
        dPc = torch.empty(g.Nn, H, device=dZ.device)
        N.check(self.lib.cb2t_edge_gather_bwd(_p(dZ), g.K, g.Nn, _p(g.rev_ptr), _p(g.rev_edge), _p(dPa), _p(dPc), self.stream), "edge_gather_bwd")
        return dPa, dPc

    def masked_sum(self, M, mask_e, g):
        S = torch.empty(g.Nn, H, device=M.device)
        N.check(self.lib.cb2t_masked_sum_fwd(_p(M), _p(mask_e), g.K, g.Nn, 1.0 / 30.0, _p(S), self.stream), "masked_sum_fwd")
        return S

    def masked_sum_bwd(self, dS, mask_e, g):
        dM = torch.empty(g.E, H, device=dS.device)
        N.check(self.lib.cb2t_masked_sum_bwd(_p(dS), _p(mask_e), g.K, g.E, 1.0 / 30.0, _p(dM), self.stream), "masked_sum_bwd")
        return dM

    def ln_mod(self, A, Bres, drop, rpm, shift, scale, gate, stride, row_mask, eps=1e-6):
        rows = A.shape[0]
        X = torch.empty_like(A) if Bres is not None else None
        stats = torch.empty(rows, 2, device=A.device)
        Y = torch.empty_like(A)
        N.check(self.lib.cb2t_ln_mod_fwd(_p(A), _p(Bres), _p(drop), rows, int(rpm), shift, scale, gate, int(stride), _p(row_mask), float(eps), _p(X),
                                         _p(stats), _p(Y), self.stream), "ln_mod_fwd")
        return Y, (X if X is not None else A), stats

    def ln_mod_bwd(self, dY, X, stats, rpm, shift, scale, gate, stride, row_mask, d_shift, d_scale, d_gate, acc=False):
        dX = torch.empty_like(X)
        N.check(self.lib.cb2t_ln_mod_bwd(_p(dY), _p(X), _p(stats), X.shape[0], int(rpm), shift, scale, gate, int(stride), _p(row_mask), _p(dX),
                                         d_shift, d_scale, d_gate, int(acc), self.stream), "ln_mod_bwd")
        return dX

    def index_sum(self, X, idx, classes, out, acc=True):
        N.check(self.lib.cb2t_index_sum(_p(X), _p(idx), X.shape[0], X.shape[1], int(classes), _p(out), int(acc), self.stream), "index_sum")


def allreduce_flat(flat_g: torch.Tensor, n_buckets: int = 4):
    """DDP semantics (train_latent.py:151-153, :251): gradients AVERAGED over the ranks, in place.  The flat buffer is reduced in
    `n_buckets` contiguous slices launched back to back (NCCL pipelines them); the layout is the reverse order of use in the forward
    pass, so the first slices are the ones the backward pass completes first."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    n = flat_g.numel()
    edges = [n * i // n_buckets for i in range(n_buckets + 1)]
    works = [dist.all_reduce(flat_g[a:b], op=dist.ReduceOp.SUM, async_op=True) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    for w in works:
        w.wait()
    flat_g.div_(world)


class Geometry:
    """Everything of a training batch that depends only on the C-alpha traces: the k-NN graph (cb2_knn_topk), per-edge neighbour node
    ids, the reverse CSR used by the gather gradients, the masks, the raw edge features (cb2t_edge_raw_features)."""

    def __init__(self, batch: dict, k_neighbors: int, device):
        num = batch["num_CGs"].to(torch.int64)
        self.B, self.L = int(num.numel()), int(num.max())
        X, cg_z = batching.pad_frames(batch["CG_nxyz"].to(device), num.to(device), self.L)
        self.X, self.cg_z = X.contiguous(), cg_z.to(torch.int32).contiguous()
        self.lengths = num.to(device, torch.int32).contiguous()
        self.K = min(int(k_neighbors), self.L)
        self.Nn, self.E = self.B * self.L, self.B * self.L * self.K
        D, idx = knn_topk(self.X, self.lengths, self.K)
        self.nbr_dist, self.nbr_idx = D, idx
        base = (torch.arange(self.B, device=device, dtype=torch.int32) * self.L).view(-1, 1, 1)
        self.nbr_node = (idx + base).reshape(-1).contiguous()                                   # [E] int32
        pos = torch.arange(self.L, device=device)
        self.mask = (pos[None, :] < num.to(device)[:, None])                                    # [B, L] bool
        self.mask_v = self.mask.reshape(-1).to(torch.float32).contiguous()                     # [N]
        mj = self.mask_v[self.nbr_node.long()].view(self.Nn, self.K)
        self.mask_e = (self.mask_v.view(-1, 1) * mj).reshape(-1).contiguous()                  # [E] mask_i * mask_j (latent_model.py:219-220)
        order = torch.argsort(self.nbr_node.long(), stable=True)
        self.rev_edge = order.to(torch.int32).contiguous()
        counts = torch.bincount(self.nbr_node.long(), minlength=self.Nn)
        self.rev_ptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=device), torch.cumsum(counts, 0)]).to(torch.int32).contiguous()
        i_idx = pos.view(1, self.L, 1).to(torch.int32)
        self.pos_class = (i_idx - idx + 32).clamp_(0, 64).reshape(-1).to(torch.int32).contiguous()   # protein_mpnn_utils.py:511-516
        self.raw = torch.empty(self.E, 152, device=device)
        N.check(N.lib().cb2t_edge_raw_features(_p(self.X), _p(idx), _p(D), self.B, self.L, self.K, _p(self.raw), N.stream_ptr()), "edge_raw_features")


class DenoiserTrainer:
    """Parameters, gradients, AdamW moments and the EMA copy live in five flat fp32 buffers; `params[name]` / `grads[name]` are views
    with the reference's state_dict names and shapes."""

    def __init__(self, state_dict: dict, k_neighbors: int = 64, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, ema_decay: float = 0.9999, grad_clip: float = 1.0, device=None, gemm: str = "fp32"):
        """gemm = "fp32": every GEMM on the fp32 SIMT path (gradient parity against fp32 autograd); "tf32": tensor cores (tcgen05.mma
        kind::tf32, fp32 accumulation) for the large GEMMs -- the arithmetic the reference trains with (train_latent.py:24-25)."""
        N.require_cuda()
        if gemm not in ("fp32", "tf32"):
            raise ValueError("gemm must be 'fp32' or 'tf32'")
        self.gemm_mode = gemm
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ops = _Ops()
        self.k_neighbors = int(k_neighbors)
        self.hp = dict(lr=lr, b1=betas[0], b2=betas[1], eps=eps, wd=weight_decay, ema=ema_decay, clip=grad_clip)
        shapes = weights.denoiser_shapes()
        # flat layout in REVERSE order of use in the forward pass, so that gradient buckets complete front to back during backward
        order = list(shapes.keys())[::-1]
        self.offsets, off = OrderedDict(), 0
        for k in order:
            n = int(math.prod(shapes[k]))
            self.offsets[k] = (off, n, tuple(shapes[k]))
            off += (n + 63) // 64 * 64                        # 256-byte aligned views (float4 loads in the GEMM)
        self.numel = off
        mk = lambda: torch.zeros(self.numel, device=self.device, dtype=torch.float32)
        self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_ema = mk(), mk(), mk(), mk(), mk()
        view = lambda flat: OrderedDict((k, flat[o:o + n].view(shp)) for k, (o, n, shp) in self.offsets.items())
        self.params, self.grads, self.ema = view(self.flat_p), view(self.flat_g), view(self.flat_ema)
        sd = {k.removeprefix("module."): v for k, v in state_dict.items()}
        for k in shapes:
            self.params[k].copy_(sd[k].to(self.device, torch.float32))
        self.flat_ema.copy_(self.flat_p)
        self._refresh_transposed()
        self.step_count = 0
        self.freqs = sinusoid_freqs().to(self.device)
        self.sumsq = torch.zeros(1, device=self.device)
        self.ctx = None

    def _refresh_transposed(self):
        """TF32 mode: W^T copies of the 2-D weights, so that the data gradients are K-contiguous on both operands (layout change only)."""
        if self.gemm_mode != "tf32":
            self.ops.transposed.clear()
            return
        for k, w in self.params.items():
            if w.dim() == 2 and w.shape[0] % 4 == 0 and w.shape[1] >= 32:
                cur = self.ops.transposed.get(w.data_ptr())
                if cur is None:
                    self.ops.transposed[w.data_ptr()] = w.t().contiguous()
                else:
                    cur.copy_(w.t())

    def _set_mode(self, on: bool = True):
        """The GEMM arithmetic is a switch of the library (cb2t_set_gemm_mode), read when a GEMM is launched: set on entry of forward /
        backward, back to the fp32 default on exit so that nothing else in the process inherits TF32."""
        N.check(self.ops.lib.cb2t_set_gemm_mode(1 if (on and self.gemm_mode == "tf32") else 0), "set_gemm_mode")
        if on:
            self.ops.stream = N.stream_ptr()

    def state_dict(self, ema: bool = False):
        src = self.ema if ema else self.params
        return OrderedDict((k, src[k].detach().clone()) for k in weights.denoiser_shapes())

    # ------------------------------------------------------------------------------------------------------------ forward
    def _adaln(self, name, c_silu, width, ctx):
        W, b = self.params[name + ".adaLN_modulation.1.weight"], self.params[name + ".adaLN_modulation.1.bias"]
        mod = self.ops.linear(c_silu, W, 0, H)
        self.ops.bias_gelu(mod, b, want_act=False)
        ctx[name + ".mod"] = mod
        return mod

    def _mlp_tail(self, A1, p, names, ctx, tag):
        """W2 -> GELU -> W3 (+ bias); keeps what the backward needs."""
        o, P = self.ops, self.params
        Z2, A2 = o.linear_bias_act(A1, P[f"{p}.{names[1]}.weight"], P[f"{p}.{names[1]}.bias"], 0, H)
        M, _ = o.linear_bias_act(A2, P[f"{p}.{names[2]}.weight"], P[f"{p}.{names[2]}.bias"], 0, H, want_act=False)
        ctx[tag] = (A1, Z2, A2)
        return M

    def _ffn(self, h, p, ctx):
        o, P = self.ops, self.params
        F1, G1 = o.linear_bias_act(h, P[f"{p}.dense.W_in.weight"], P[f"{p}.dense.W_in.bias"], 0, H)
        F2, _ = o.linear_bias_act(G1, P[f"{p}.dense.W_out.weight"], P[f"{p}.dense.W_out.bias"], 0, 4 * H, want_act=False)
        ctx[p + ".ffn"] = (h, F1, G1)
        return F2

    def _drop(self, like, p, gen):
        if p <= 0.0:
            return None
        keep = torch.rand(like.shape, device=like.device, generator=gen) >= p
        return keep.to(torch.float32).div_(1.0 - p)

    def _enc_layer(self, l, hV, hE, c_silu, g, ctx, p_drop, gen):
        o, P = self.ops, self.params
        p = f"encoder_layers.{l}"
        mod = self._adaln(p, c_silu, 9 * H, ctx)
        mp, ms = mod.data_ptr(), mod.stride(0)
        chunk = lambda q: mp + 4 * H * q
        # node message: W1 [128, 384] = [h_V_i | h_E | h_V_j]
        W1 = P[p + ".W1.weight"]
        Pa, Pc = o.linear(hV, W1, 0, H), o.linear(hV, W1, 2 * H, H)
        Z1 = o.linear(hE, W1, H, H)
        A1 = o.edge_combine_gelu(Z1, Pa, Pc, P[p + ".W1.bias"], g)
        M = self._mlp_tail(A1, p, ("W1", "W2", "W3"), ctx, p + ".msg")
        dh = o.masked_sum(M, g.mask_e, g)
        d1 = self._drop(dh, p_drop, gen)
        hV1, X1, st1 = o.ln_mod(hV, dh, d1, g.L, chunk(0), chunk(1), chunk(2), ms, None)
        F2 = self._ffn(hV1, p, ctx)
        d2 = self._drop(F2, p_drop, gen)
        hV2, X2, st2 = o.ln_mod(hV1, F2, d2, g.L, chunk(3), chunk(4), chunk(5), ms, g.mask_v)
        # edge update: W11 [128, 384]
        W11 = P[p + ".W11.weight"]
        Pa2, Pc2 = o.linear(hV2, W11, 0, H), o.linear(hV2, W11, 2 * H, H)
        Z11 = o.linear(hE, W11, H, H)
        A11 = o.edge_combine_gelu(Z11, Pa2, Pc2, P[p + ".W11.bias"], g)
        M3 = self._mlp_tail(A11, p, ("W11", "W12", "W13"), ctx, p + ".upd")
        d3 = self._drop(M3, p_drop, gen)
        hE2, X3, st3 = o.ln_mod(hE, M3, d3, g.L * g.K, chunk(6), chunk(7), chunk(8), ms, None)
        ctx[p] = dict(hV=hV, hE=hE, Z1=Z1, Z11=Z11, X1=X1, st1=st1, X2=X2, st2=st2, X3=X3, st3=st3, hV2=hV2, d1=d1, d2=d2, d3=d3)
        return hV2, hE2

    def _dec_layer(self, l, hV, hE2x, hS2, hVenc, c_silu, g, ctx, p_drop, gen):
        o, P = self.ops, self.params
        p = f"decoder_layers.{l}"
        mod = self._adaln(p, c_silu, 6 * H, ctx)
        mp, ms = mod.data_ptr(), mod.stride(0)
        chunk = lambda q: mp + 4 * H * q
        # W1 [128, 512] over [h_V_i | 2 h_E | 2 h_S_j | h_V_j + h_Venc_j]  (latent_model.py:258-262)
        W1 = P[p + ".W1.weight"]
        hsum = o.ew(2, hV, hVenc)
        Pa = o.linear(hV, W1, 0, H)
        Pc = o.linear(hS2, W1, 2 * H, H)
        o.linear(hsum, W1, 3 * H, H, out=Pc, acc=True)
        Z1 = o.linear(hE2x, W1, H, H)
        A1 = o.edge_combine_gelu(Z1, Pa, Pc, P[p + ".W1.bias"], g)
        M = self._mlp_tail(A1, p, ("W1", "W2", "W3"), ctx, p + ".msg")
        dh = o.masked_sum(M, None, g)                              # the decoder passes no neighbour mask
        d1 = self._drop(dh, p_drop, gen)
        hV1, X1, st1 = o.ln_mod(hV, dh, d1, g.L, chunk(0), chunk(1), chunk(2), ms, None)
        F2 = self._ffn(hV1, p, ctx)
        d2 = self._drop(F2, p_drop, gen)
        hV2, X2, st2 = o.ln_mod(hV1, F2, d2, g.L, chunk(3), chunk(4), chunk(5), ms, g.mask_v)
        ctx[p] = dict(hV=hV, hsum=hsum, Z1=Z1, X1=X1, st1=st1, X2=X2, st2=st2, d1=d1, d2=d2)
        return hV2

    def edge_features(self, g: Geometry, ctx: dict):
        """CA_ProteinFeatures' trainable half + W_e (protein_mpnn_utils.py:511-523, latent_model.py:216) on the raw features of `g`:
        positional table through edge_embedding[:, :16], raw features through [:, 16:], affine LayerNorm, W_e -> h_E0 [E, 128]."""
        o, P = self.ops, self.params
        Wp, bp, We = P["features.embeddings.linear.weight"], P["features.embeddings.linear.bias"], P["features.edge_embedding.weight"]
        posT = (Wp[:, :65].t() + bp[None, :]).contiguous()                                # [65, 16]  (class 65 is never produced)
        PT = o.linear(posT, We, 0, 16)                                                    # [65, 128]
        Epre = o.linear(g.raw, We, 16, 151)
        N.check(o.lib.cb2t_row_gather_add(_p(Epre), _p(PT), _p(g.pos_class), g.E, N.stream_ptr()), "row_gather_add")
        lnw, lnb = P["features.norm_edges.weight"], P["features.norm_edges.bias"]
        lnw1 = (lnw - 1.0).contiguous()
        Efeat, _, stE = o.ln_mod(Epre, None, None, g.E, lnb.data_ptr(), lnw1.data_ptr(), None, 0, None, eps=1e-5)
        hE, _ = o.linear_bias_act(Efeat, P["W_e.weight"], P["W_e.bias"], 0, H, want_act=False)
        ctx["feat"] = (posT, Epre, stE, Efeat, lnw1)
        return hE

    def forward(self, x, t, geom: Geometry, dropout_p: float = 0.0, generator=None, hE0=None, hS2=None, keep: bool = True):
        """x [B, L, 3] fp32, t [B] (original 0..999 scale) -> model output [B, L, 6].  Training: the activations the backward needs are
        kept (keep=True).  Inference (tf32_tier.py): hE0 / hS2 are the per-frame quantities hoisted out of the step loop, keep=False."""
        o, P, g = self.ops, self.params, geom
        self._set_mode()
        ctx = {"geom": g}
        x2 = x.to(self.device, torch.float32).reshape(g.Nn, 3).contiguous()
        # timestep embedder (latent_model.py:37-75) and the SiLU in front of every adaLN projection
        args = t.to(self.device, torch.float32)[:, None] * self.freqs[None]
        tf = torch.cat([torch.cos(args), torch.sin(args)], dim=-1).contiguous()          # input featurisation (no parameters)
        T0 = o.linear(tf, P["t_embedder.mlp.0.weight"], 0, 256)
        o.bias_gelu(T0, P["t_embedder.mlp.0.bias"], want_act=False)
        T0a = o.ew(0, T0)
        c = o.linear(T0a, P["t_embedder.mlp.2.weight"], 0, H)
        o.bias_gelu(c, P["t_embedder.mlp.2.bias"], want_act=False)
        c_silu = o.ew(0, c)
        ctx["temb"] = (tf, T0, T0a, c, c_silu)
        hE = hE0 if hE0 is not None else self.edge_features(g, ctx)
        hV = o.linear(x2, P["x_in.weight"], 0, 3)
        o.bias_gelu(hV, P["x_in.bias"], want_act=False)
        ctx["x2"] = x2
        for l in range(3):
            hV, hE = self._enc_layer(l, hV, hE, c_silu, g, ctx, dropout_p, generator)
        hVenc = hV
        if hS2 is None:
            hS = P["W_s.weight"][g.cg_z.reshape(-1).long()]                               # embedding lookup (index plumbing), latent_model.py:225
            hS2 = o.ew(3, hS.contiguous(), scale=2.0)
        hE2x = o.ew(3, hE, scale=2.0)
        ctx["dec_in"] = (hE, hS2, hE2x, hVenc)
        for l in range(3):
            hV = self._dec_layer(l, hV, hE2x, hS2, hVenc, c_silu, g, ctx, dropout_p, generator)
        # FinalLayer (latent_model.py:21-35)
        modf = self._adaln("W_out", c_silu, 2 * H, ctx)
        Yf, _, stF = o.ln_mod(hV, None, None, g.L, modf.data_ptr(), modf.data_ptr() + 4 * H, None, modf.stride(0), None)
        out = o.linear(Yf, P["W_out.linear.weight"], 0, H)
        o.bias_gelu(out, P["W_out.linear.bias"], want_act=False)
        ctx["final"] = (hV, stF, Yf)
        self.ctx = ctx if keep else None
        self._set_mode(False)
        return out.view(g.B, g.L, 6)

    # ------------------------------------------------------------------------------------------------------------ backward
    def _lin_bwd(self, dy, x, wname, c0, kin, need_dx=True, bias=True, dx_out=None, dx_acc=False):
        o, P, G = self.ops, self.params, self.grads
        o.linear_dw(dy, x, G[wname + ".weight"], c0)
        if bias:
            o.colsum(dy, G[wname + ".bias"])
        if need_dx:
            return o.linear_dx(dy, P[wname + ".weight"], c0, kin, out=dx_out, acc=dx_acc)
        return None

    def _mlp_tail_bwd(self, dM, p, names, tag):
        """Backward of W3(GELU(W2 A1 + b2)) + b3 -> dA1."""
        A1, Z2, A2 = self.ctx[tag]
        dA2 = self._lin_bwd(dM, A2, f"{p}.{names[2]}", 0, H)
        dZ2 = self.ops.gelu_bwd(Z2, dA2, bias_grad=self.grads[f"{p}.{names[1]}.bias"])
        return self._lin_bwd(dZ2, A1, f"{p}.{names[1]}", 0, H, bias=False)

    def _ffn_bwd(self, dF2, p, dh_acc):
        h, F1, G1 = self.ctx[p + ".ffn"]
        dG1 = self._lin_bwd(dF2, G1, p + ".dense.W_out", 0, 4 * H)
        dF1 = self.ops.gelu_bwd(F1, dG1, bias_grad=self.grads[p + ".dense.W_in.bias"])
        self._lin_bwd(dF1, h, p + ".dense.W_in", 0, H, dx_out=dh_acc, dx_acc=True, bias=False)

    def _adaln_bwd(self, name, dmod, d_c_silu):
        c_silu = self.ctx["temb"][4]
        self._lin_bwd(dmod, c_silu, name + ".adaLN_modulation.1", 0, H, dx_out=d_c_silu, dx_acc=True)

    def _mul(self, a, d):
        return a if d is None else self.ops.ew(4, a, d)

    def _enc_layer_bwd(self, l, dhV2, dhE2, d_c_silu):
        o, P, G, g = self.ops, self.params, self.grads, self.ctx["geom"]
        p = f"encoder_layers.{l}"
        s = self.ctx[p]
        mod = self.ctx[p + ".mod"]
        dmod = torch.zeros_like(mod)
        mp, dp, ms = mod.data_ptr(), dmod.data_ptr(), mod.stride(0)
        ch, dch = (lambda q: mp + 4 * H * q), (lambda q: dp + 4 * H * q)
        # ---- edge update
        dX3 = o.ln_mod_bwd(dhE2, s["X3"], s["st3"], g.L * g.K, ch(6), ch(7), ch(8), ms, None, dch(6), dch(7), dch(8))
        dhE = dX3                                                                          # residual branch (aliased on purpose, read-only from here)
        dM3 = self._mul(dX3, s["d3"])
        dA11 = self._mlp_tail_bwd(dM3, p, ("W11", "W12", "W13"), p + ".upd")
        dZ11 = o.gelu_bwd(s["Z11"], dA11, bias_grad=G[p + ".W11.bias"])
        W11 = P[p + ".W11.weight"]
        o.linear_dw(dZ11, s["hE"], G[p + ".W11.weight"], H)
        dhE_total = o.linear_dx(dZ11, W11, H, H)
        o.ew(2, dhE_total, dhE, out=dhE_total)
        dPa2, dPc2 = o.edge_gather_bwd(dZ11, g)
        o.linear_dw(dPa2, s["hV2"], G[p + ".W11.weight"], 0)
        o.linear_dw(dPc2, s["hV2"], G[p + ".W11.weight"], 2 * H)
        dhV2 = dhV2.clone() if dhV2 is not None else torch.zeros(g.Nn, H, device=self.device)
        o.linear_dx(dPa2, W11, 0, H, out=dhV2, acc=True)
        o.linear_dx(dPc2, W11, 2 * H, H, out=dhV2, acc=True)
        # ---- node update
        dX2 = o.ln_mod_bwd(dhV2, s["X2"], s["st2"], g.L, ch(3), ch(4), ch(5), ms, g.mask_v, dch(3), dch(4), dch(5))
        dhV1 = dX2.clone()
        self._ffn_bwd(self._mul(dX2, s["d2"]), p, dhV1)
        dX1 = o.ln_mod_bwd(dhV1, s["X1"], s["st1"], g.L, ch(0), ch(1), ch(2), ms, None, dch(0), dch(1), dch(2))
        dhV = dX1.clone()
        dM = o.masked_sum_bwd(self._mul(dX1, s["d1"]), g.mask_e, g)
        dA1 = self._mlp_tail_bwd(dM, p, ("W1", "W2", "W3"), p + ".msg")
        dZ1 = o.gelu_bwd(s["Z1"], dA1, bias_grad=G[p + ".W1.bias"])
        W1 = P[p + ".W1.weight"]
        o.linear_dw(dZ1, s["hE"], G[p + ".W1.weight"], H)
        o.linear_dx(dZ1, W1, H, H, out=dhE_total, acc=True)
        dPa, dPc = o.edge_gather_bwd(dZ1, g)
        o.linear_dw(dPa, s["hV"], G[p + ".W1.weight"], 0)
        o.linear_dw(dPc, s["hV"], G[p + ".W1.weight"], 2 * H)
        o.linear_dx(dPa, W1, 0, H, out=dhV, acc=True)
        o.linear_dx(dPc, W1, 2 * H, H, out=dhV, acc=True)
        self._adaln_bwd(p, dmod, d_c_silu)
        return dhV, dhE_total

    def _dec_layer_bwd(self, l, dhV2, dhE2x, dhS2, dhVenc, d_c_silu):
        o, P, G, g = self.ops, self.params, self.grads, self.ctx["geom"]
        p = f"decoder_layers.{l}"
        s = self.ctx[p]
        hE, hS2, hE2x, hVenc = self.ctx["dec_in"]
        mod = self.ctx[p + ".mod"]
        dmod = torch.zeros_like(mod)
        mp, dp, ms = mod.data_ptr(), dmod.data_ptr(), mod.stride(0)
        ch, dch = (lambda q: mp + 4 * H * q), (lambda q: dp + 4 * H * q)
        dX2 = o.ln_mod_bwd(dhV2, s["X2"], s["st2"], g.L, ch(3), ch(4), ch(5), ms, g.mask_v, dch(3), dch(4), dch(5))
        dhV1 = dX2.clone()
        self._ffn_bwd(self._mul(dX2, s["d2"]), p, dhV1)
        dX1 = o.ln_mod_bwd(dhV1, s["X1"], s["st1"], g.L, ch(0), ch(1), ch(2), ms, None, dch(0), dch(1), dch(2))
        dhV = dX1.clone()
        dM = o.masked_sum_bwd(self._mul(dX1, s["d1"]), None, g)
        dA1 = self._mlp_tail_bwd(dM, p, ("W1", "W2", "W3"), p + ".msg")
        dZ1 = o.gelu_bwd(s["Z1"], dA1, bias_grad=G[p + ".W1.bias"])
        W1 = P[p + ".W1.weight"]
        o.linear_dw(dZ1, hE2x, G[p + ".W1.weight"], H)
        o.linear_dx(dZ1, W1, H, H, out=dhE2x, acc=True)
        dPa, dPc = o.edge_gather_bwd(dZ1, g)
        o.linear_dw(dPa, s["hV"], G[p + ".W1.weight"], 0)
        o.linear_dx(dPa, W1, 0, H, out=dhV, acc=True)
        o.linear_dw(dPc, hS2, G[p + ".W1.weight"], 2 * H)
        o.linear_dx(dPc, W1, 2 * H, H, out=dhS2, acc=True)
        o.linear_dw(dPc, s["hsum"], G[p + ".W1.weight"], 3 * H)
        dsum = o.linear_dx(dPc, W1, 3 * H, H)
        o.ew(2, dhV, dsum, out=dhV)
        o.ew(2, dhVenc, dsum, out=dhVenc)
        self._adaln_bwd(p, dmod, d_c_silu)
        return dhV

    def backward(self, dout):
        """dout [B, L, 6] = d loss / d model output -> self.grads (+=; call zero_grad() first)."""
        o, P, G, ctx = self.ops, self.params, self.grads, self.ctx
        self._set_mode()
        g = ctx["geom"]
        d_c_silu = torch.zeros(g.B, H, device=self.device)
        dout2 = dout.to(self.device, torch.float32).reshape(g.Nn, 6).contiguous()
        hVf, stF, Yf = ctx["final"]
        dYf = self._lin_bwd(dout2, Yf, "W_out.linear", 0, H)
        modf = ctx["W_out.mod"]
        dmodf = torch.zeros_like(modf)
        dhV = o.ln_mod_bwd(dYf, hVf, stF, g.L, modf.data_ptr(), modf.data_ptr() + 4 * H, None, modf.stride(0), None,
                           dmodf.data_ptr(), dmodf.data_ptr() + 4 * H, None)
        self._adaln_bwd("W_out", dmodf, d_c_silu)
        hE, hS2, hE2x, hVenc = ctx["dec_in"]
        dhE2x = torch.zeros_like(hE2x)
        dhS2 = torch.zeros_like(hS2)
        dhVenc = torch.zeros_like(hVenc)
        for l in (2, 1, 0):
            dhV = self._dec_layer_bwd(l, dhV, dhE2x, dhS2, dhVenc, d_c_silu)
        o.ew(2, dhV, dhVenc, out=dhV)                                                       # decoder layer 0 starts from the encoder output
        o.index_sum(o.ew(3, dhS2, scale=2.0), g.cg_z.reshape(-1), P["W_s.weight"].shape[0], G["W_s.weight"])
        dhE = o.ew(3, dhE2x, scale=2.0)
        for l in (2, 1, 0):
            dhV, dhE = self._enc_layer_bwd(l, dhV, dhE, d_c_silu)
        self._lin_bwd(dhV, ctx["x2"], "x_in", 0, 3, need_dx=False)
        # featuriser parameters
        posT, Epre, stE, Efeat, lnw1 = ctx["feat"]
        dEfeat = self._lin_bwd(dhE, Efeat, "W_e", 0, H)
        dlnw = torch.zeros(1, H, device=self.device)
        dlnb = torch.zeros(1, H, device=self.device)
        dEpre = o.ln_mod_bwd(dEfeat, Epre, stE, g.E, P["features.norm_edges.bias"].data_ptr(), lnw1.data_ptr(), None, 0, None,
                             dlnb.data_ptr(), dlnw.data_ptr(), None)
        G["features.norm_edges.weight"].add_(dlnw[0])
        G["features.norm_edges.bias"].add_(dlnb[0])
        We, dWe = P["features.edge_embedding.weight"], G["features.edge_embedding.weight"]
        o.linear_dw(dEpre, g.raw[:, :151], dWe, 16)
        dPT = torch.zeros(65, H, device=self.device)
        o.index_sum(dEpre, g.pos_class, 65, dPT, acc=False)
        o.linear_dw(dPT, posT, dWe, 0)
        dposT = o.linear_dx(dPT, We, 0, 16)                                                 # [65, 16]
        G["features.embeddings.linear.weight"][:, :65].add_(dposT.t())
        G["features.embeddings.linear.bias"].add_(dposT.sum(0))
        # timestep embedder
        tf, T0, T0a, c, c_silu = ctx["temb"]
        dc = o.ew(1, c, d_c_silu)
        dT0a = self._lin_bwd(dc, T0a, "t_embedder.mlp.2", 0, H)
        dT0 = o.ew(1, T0, dT0a)
        self._lin_bwd(dT0, tf, "t_embedder.mlp.0", 0, 256, need_dx=False)
        self.ctx = None
        self._set_mode(False)

    # ------------------------------------------------------------------------------------------------------------ optimiser
    def zero_grad(self):
        self.flat_g.zero_()

    def allreduce_grads(self, n_buckets: int = 4):
        allreduce_flat(self.flat_g, n_buckets)

    def step(self, lr_scale: float = 1.0):
        """clip_grad_norm_(max_norm) + AdamW + EMA in two kernels (sum of squares, then the fused update); no host synchronisation."""
        hp = self.hp
        self.step_count += 1
        lib = self.ops.lib
        N.check(lib.cb2t_sumsq(_p(self.flat_g), self.numel, _p(self.sumsq), N.stream_ptr()), "sumsq")
        N.check(lib.cb2t_adamw_ema(_p(self.flat_p), _p(self.flat_g), _p(self.flat_m), _p(self.flat_v), _p(self.flat_ema), self.numel,
                                   hp["lr"] * lr_scale, hp["b1"], hp["b2"], hp["eps"], hp["wd"], self.step_count, hp["ema"],
                                   _p(self.sumsq), hp["clip"] if hp["clip"] else 0.0, N.stream_ptr()), "adamw_ema")
        self._refresh_transposed()

    def grad_norm(self) -> float:
        return float(self.sumsq.sqrt().item())

    # ------------------------------------------------------------------------------------------------------------ one training step
    def train_step(self, diffusion, x1, t, batch: dict, noise=None, dropout_p: float = 0.0, generator=None, geom: Geometry = None,
                   do_step: bool = True, zero: bool = True, loss_weight: float = 1.0, do_allreduce: bool = True):
        """train_latent.py:203-261 for one batch: x1 [B, L, 3] normalised latents, t [B] in [0, T).  Returns the loss terms.
        Gradient accumulation over micro-batches: zero=False keeps the gradients of the previous call, loss_weight scales this
        call's share of the batch mean (micro-batch size / batch size), do_allreduce / do_step only on the last micro-batch."""
        g = geom if geom is not None else Geometry(batch, self.k_neighbors, self.device)
        x1 = x1.to(self.device, torch.float32)
        t = t.to(self.device)
        noise = torch.randn_like(x1) if noise is None else noise.to(self.device, torch.float32)
        x_t = diffusion.q_sample(x1, t, noise)
        map_t = torch.tensor(diffusion.timestep_map, device=self.device, dtype=t.dtype)[t]
        if zero:
            self.zero_grad()
        out = self.forward(x_t, map_t, g, dropout_p, generator)
        leaf = out.detach().requires_grad_(True)
        with torch.enable_grad():                       # the scalar loss on the [B, L, 6] output (gaussian_diffusion.py:598-725)
            terms = diffusion.training_losses(lambda *a, **k: leaf, x1, t, dict(mask=g.mask), noise=noise)
            loss = terms["loss"].mean()
            (dout,) = torch.autograd.grad(loss * loss_weight, leaf)
        self.backward(dout)
        if do_allreduce:
            self.allreduce_grads()
        if do_step:
            self.step()
        return {k: v.detach() for k, v in terms.items()}, loss.detach()          # 0-d device tensor: no host synchronisation inside a step
