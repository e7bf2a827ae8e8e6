"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A CPU (torch fp32 / numpy float64) restatement of the reference's reverse latent-diffusion
sampling path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this module; the product path (`codlad_b200/`) never does
and fails loudly when the CUDA library is missing.

Every function cites the reference file:line it restates (paths relative to the reference
checkout).  The restatement is functional (weights come in as a `state_dict`-style mapping with
the reference's own key names) and edge-centric (features are computed per (i, k) neighbour
pair, never as dense [L, L] tensors), which is also how the CUDA kernels are organised; it is
pinned against the reference itself by `oracle/make_goldens.py` + `tests/test_oracle_golden.py`
(goldens produced by running the unmodified reference modules in the build container).

Parity status:
  * denoiser / diffusion / IC decoder / ic_to_xyz: PINNED (goldens generated from the reference).
  * VectorQuantize eval path: the arithmetic lives in the un-vendored third-party package
    vector_quantize_pytorch==1.21.7 (requirements.txt:34) -> restated from its published
    algorithm; cross-checked against the in-repo equivalent utils/vq_module.py:56-71
    (VectorQuantizerEMA).  PARITY UNPINNED for near-tie rows and for the index value reported
    at masked positions.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

H = 128  # hidden width of the denoiser (models/latent_model.py:80-84)


def sqrt_rn(t: torch.Tensor) -> torch.Tensor:
    """IEEE correctly-rounded fp32 square root (numpy -> hardware sqrtps).  torch.sqrt on the
    AVX512 CPU build of torch 2.11 is NOT correctly rounded (1 ulp low on ~0.6 % of inputs,
    measured against float64 and against torch.sqrt on CUDA, which is correctly rounded), so the
    places where this oracle promises bit-exact results (k-NN distances, VQ distances) use this
    helper: it is what the reference computes on its own target (a CUDA device)."""
    return torch.from_numpy(np.sqrt(t.detach().contiguous().numpy()))


# --------------------------------------------------------------------------------------
# Diffusion schedule (float64 numpy, like the reference)
# --------------------------------------------------------------------------------------
def kept_timesteps(num_steps: int, diffusion_steps: int = 1000):
    """diffusion_and_flow/respace.py:12-62 for a single section given as str(num_steps)."""
    if num_steps <= 1:
        stride = 1.0
    else:
        stride = (diffusion_steps - 1) / (num_steps - 1)
    cur, out = 0.0, []
    for _ in range(num_steps):
        out.append(round(cur))
        cur += stride
    return sorted(set(out))


def respaced_schedule(num_steps: int = 100, diffusion_steps: int = 1000):
    """Linear betas (gaussian_diffusion.py:104-121), respacing (respace.py:73-87) and the
    coefficient tables (gaussian_diffusion.py:175-209).  Returns float64 arrays."""
    scale = 1000 / diffusion_steps
    base_betas = np.linspace(scale * 0.0001, scale * 0.02, diffusion_steps, dtype=np.float64)
    base_ac = np.cumprod(1.0 - base_betas, axis=0)
    keep = set(kept_timesteps(num_steps, diffusion_steps))
    last, betas, tmap = 1.0, [], []
    for i, ac in enumerate(base_ac):
        if i in keep:
            betas.append(1 - ac / last)
            last = ac
            tmap.append(i)
    betas = np.array(betas, dtype=np.float64)
    ac = np.cumprod(1.0 - betas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    return {
        "timestep_map": np.array(tmap, dtype=np.int64),
        "betas": betas,
        "log_betas": np.log(betas),
        "sqrt_recip_ac": np.sqrt(1.0 / ac),
        "sqrt_recipm1_ac": np.sqrt(1.0 / ac - 1),
        "post_logvar_clipped": np.log(np.append(post_var[1], post_var[1:])),
        "post_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "post_coef2": (1.0 - ac_prev) * np.sqrt(1.0 - betas) / (1.0 - ac),
    }


def p_sample_update(x, model_out, step, noise, sched):
    """gaussian_diffusion.py:303-318 (learned-range variance), :345-351 (x0 from eps, posterior
    mean), :440-446 (sample).  `step` is the respaced index (0..T-1), scalar."""
    C = x.shape[-1]
    eps, v = model_out[..., :C], model_out[..., C:]
    f32 = lambda a: torch.tensor(float(np.float32(a[step])), dtype=torch.float32)
    min_log, max_log = f32(sched["post_logvar_clipped"]), f32(sched["log_betas"])
    frac = (v + 1) / 2
    logvar = frac * max_log + (1 - frac) * min_log
    x0 = f32(sched["sqrt_recip_ac"]) * x - f32(sched["sqrt_recipm1_ac"]) * eps
    mean = f32(sched["post_coef1"]) * x0 + f32(sched["post_coef2"]) * x
    nz = 0.0 if step == 0 else 1.0
    return mean + nz * torch.exp(0.5 * logvar) * noise


# --------------------------------------------------------------------------------------
# Denoiser pieces
# --------------------------------------------------------------------------------------
def lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def timestep_cond(sd, t):
    """models/latent_model.py:37-75: sinusoid(256) -> Linear -> SiLU -> Linear; t is [B]."""
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    return lin(sd, "t_embedder.mlp.2", F.silu(lin(sd, "t_embedder.mlp.0", emb)))


def knn_graph(X, mask, k_neighbors):
    """models/protein_mpnn_utils.py:447-459.  X [B,L,3] fp32, mask [B,L] {0,1}.
    Distances are formed exactly as ((dx*dx + dy*dy) + dz*dz) + 1e-6 then a correctly rounded
    sqrt, all in fp32 without FMA contraction.  Selection is the K smallest of D_adjust ordered
    by (distance, index) -- the reference's torch.topk leaves the order of exactly tied
    distances unspecified, so parity against it is tie-aware (tests/parity_utils.py).
    Returns D_neighbors [B,L,K] fp32, E_idx [B,L,K] int64."""
    B, L, _ = X.shape
    K = min(int(k_neighbors), L)
    m = mask.to(torch.float32)
    d = X[:, None, :, :] - X[:, :, None, :]          # d[b,i,j] = X[b,j] - X[b,i]
    sq = d * d
    s = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    m2 = m[:, None, :] * m[:, :, None]
    D = m2 * sqrt_rn(s + 1e-6)
    Dmax = D.max(-1, keepdim=True).values
    Dadj = D + (1.0 - m2) * Dmax
    vals, idx = torch.sort(Dadj, dim=-1, stable=True)
    return vals[..., :K].contiguous(), idx[..., :K].contiguous()


def _unit(v, eps=1e-12):
    return v / v.norm(dim=-1, keepdim=True).clamp_min(eps)  # F.normalize


def residue_frames(X):
    """models/protein_mpnn_utils.py:397-426: local frames O_i = rows (o1, n2, o1 x n2) for
    1 <= i <= L-3 of the *padded* length, zero elsewhere.  Returns [B,L,3,3]."""
    B, L, _ = X.shape
    dX = X[:, 1:] - X[:, :-1]
    nrm = torch.norm(dX, dim=-1)
    gate = (3.6 < nrm) & (nrm < 4.0)
    U = _unit(dX * gate[..., None])
    u2, u1 = U[:, :-2], U[:, 1:-1]
    n2 = _unit(torch.linalg.cross(u2, u1))
    o1 = _unit(u2 - u1)
    O = torch.stack((o1, n2, torch.linalg.cross(o1, n2)), dim=2)  # [B,L-3,3,3]
    out = torch.zeros(B, L, 3, 3, dtype=X.dtype)
    if L > 3:
        out[:, 1:L - 2] = O
    return out


def _gather_nodes(nodes, E_idx):
    """nodes [B,L,...] gathered at E_idx [B,L,K] -> [B,L,K,...] (protein_mpnn_utils.py:103-111)."""
    B, L, K = E_idx.shape
    flat = E_idx.reshape(B, L * K)
    tail = nodes.shape[2:]
    idx = flat.reshape(B, L * K, *([1] * len(tail))).expand(B, L * K, *tail)
    return torch.gather(nodes, 1, idx).reshape(B, L, K, *tail)


RBF_MU = torch.linspace(2.0, 22.0, 16)
RBF_SIGMA = (22.0 - 2.0) / 16


def _rbf16(D):
    """protein_mpnn_utils.py:461-470."""
    return torch.exp(-(((D[..., None] - RBF_MU) / RBF_SIGMA) ** 2))


def edge_input_features(X, D_nb, E_idx):
    """The 167-wide raw edge feature cat(pos-class one-hot handled separately, RBF144, O7).
    protein_mpnn_utils.py:478-517.  Returns (pos_class [B,L,K] int64, rbf [B,L,K,144], ori [B,L,K,7])."""
    B, L, K = E_idx.shape
    prev = torch.zeros_like(X)
    nxt = torch.zeros_like(X)
    prev[:, 1:] = X[:, :-1]
    nxt[:, :-1] = X[:, 1:]
    trio_i = (prev, X, nxt)
    trio_j = tuple(_gather_nodes(a, E_idx) for a in trio_i)

    def pair(a, b):
        d = trio_i[a][:, :, None, :] - trio_j[b]
        return torch.sqrt((d * d).sum(-1) + 1e-6)

    rbf = [_rbf16(D_nb)]
    for a, b in ((0, 0), (2, 2), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)):
        rbf.append(_rbf16(pair(a, b)))
    rbf = torch.cat(rbf, dim=-1)

    O = residue_frames(X)
    O_j = _gather_nodes(O, E_idx)                                  # [B,L,K,3,3]
    dXij = trio_j[1] - X[:, :, None, :]
    dU = _unit(torch.einsum("blrc,blkc->blkr", O, dXij))
    R = torch.einsum("blra,blkrc->blkac", O, O_j)                   # O_i^T O_j
    q = quaternion_of(R)
    ori = torch.cat((dU, q), dim=-1)

    ar = torch.arange(L)
    pos = (ar[None, :, None] - E_idx + 32).clamp(0, 64)
    return pos, rbf, ori


def quaternion_of(R):
    """protein_mpnn_utils.py:369-395."""
    Rxx, Ryy, Rzz = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    mag = 0.5 * torch.sqrt(torch.abs(1 + torch.stack([Rxx - Ryy - Rzz, -Rxx + Ryy - Rzz, -Rxx - Ryy + Rzz], -1)))
    sgn = torch.sign(torch.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], -1))
    w = torch.sqrt(F.relu(1 + Rxx + Ryy + Rzz))[..., None] / 2.0
    return _unit(torch.cat((sgn * mag, w), -1))


def edge_embedding(sd, X, mask, k_neighbors):
    """CA_ProteinFeatures.forward (protein_mpnn_utils.py:478-523) + W_e (latent_model.py:216).
    Returns E_idx, D_nb, E (after norm_edges), h_E0 = W_e(E)."""
    D_nb, E_idx = knn_graph(X, mask, k_neighbors)
    pos, rbf, ori = edge_input_features(X, D_nb, E_idx)
    Wp, bp = sd["features.embeddings.linear.weight"], sd["features.embeddings.linear.bias"]
    e_pos = Wp.t()[pos] + bp                                       # one-hot(66) @ W^T + b
    raw = torch.cat((e_pos, rbf, ori), dim=-1)
    E = F.linear(raw, sd["features.edge_embedding.weight"])
    E = F.layer_norm(E, (H,), sd["features.norm_edges.weight"], sd["features.norm_edges.bias"], 1e-5)
    return E_idx, D_nb, E, lin(sd, "W_e", E)


def _ln(x):
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)


def _mod(x, shift, scale):
    return x * (1 + scale) + shift


def _ffn(sd, p, h):
    return lin(sd, p + ".dense.W_out", F.gelu(lin(sd, p + ".dense.W_in", h)))


def _mlp3(sd, p, names, x):
    a, b, c = names
    return lin(sd, f"{p}.{c}", F.gelu(lin(sd, f"{p}.{b}", F.gelu(lin(sd, f"{p}.{a}", x)))))


def encoder_layer(sd, p, h_V, h_E, E_idx, mask, mask_attend, c):
    """EncLayer_diffusion.forward, protein_mpnn_utils.py:236-271 (eval mode: dropout = identity)."""
    K = E_idx.shape[-1]
    m = lin(sd, p + ".adaLN_modulation.1", F.silu(c))[:, None, :].chunk(9, dim=-1)
    cat = torch.cat([h_V[:, :, None, :].expand(-1, -1, K, -1), h_E, _gather_nodes(h_V, E_idx)], -1)
    msg = _mlp3(sd, p, ("W1", "W2", "W3"), cat) * mask_attend[..., None]
    h_V = m[2] * _mod(_ln(h_V + msg.sum(-2) / 30.0), m[0], m[1])
    h_V = m[5] * _mod(_ln(h_V + _ffn(sd, p, h_V)), m[3], m[4])
    h_V = mask[..., None] * h_V
    cat = torch.cat([h_V[:, :, None, :].expand(-1, -1, K, -1), h_E, _gather_nodes(h_V, E_idx)], -1)
    msg = _mlp3(sd, p, ("W11", "W12", "W13"), cat)
    h_E = m[8][:, :, None, :] * _mod(_ln(h_E + msg), m[6][:, :, None, :], m[7][:, :, None, :])
    return h_V, h_E


def decoder_layer(sd, p, h_V, h_ESV, mask, c):
    """DecLayer_diffusion.forward, protein_mpnn_utils.py:296-318 (mask_attend=None)."""
    K = h_ESV.shape[-2]
    m = lin(sd, p + ".adaLN_modulation.1", F.silu(c))[:, None, :].chunk(6, dim=-1)
    cat = torch.cat([h_V[:, :, None, :].expand(-1, -1, K, -1), h_ESV], -1)
    msg = _mlp3(sd, p, ("W1", "W2", "W3"), cat)
    h_V = m[2] * _mod(_ln(h_V + msg.sum(-2) / 30.0), m[0], m[1])
    h_V = m[5] * _mod(_ln(h_V + _ffn(sd, p, h_V)), m[3], m[4])
    return mask[..., None] * h_V


def denoiser_forward(sd, x, t, X, cg_z, mask, k_neighbors=64, graph=None, trace=None):
    """ProteinMPNN_diffusion_new.forward, models/latent_model.py:175-268, for the mpnn_diffusion
    factory (:276-281: decoder_mask=False, use_seq_in_encoder=True, augment_eps=0).
    x [B,L,3], t [B] (original-scale timestep), X [B,L,3] padded C-alpha, cg_z [B,L] int64,
    mask [B,L] bool.  `graph` may carry a precomputed (E_idx, h_E0) pair (the features depend
    only on X).  Returns eps/var logits [B,L,6]."""
    maskf = mask.to(torch.float32)
    c = timestep_cond(sd, t)
    if graph is None:
        E_idx, _, _, h_E = edge_embedding(sd, X, mask.to(torch.int32), k_neighbors)
    else:
        E_idx, h_E = graph
    h_V = lin(sd, "x_in", x)
    mask_attend = maskf[..., None] * _gather_nodes(maskf[..., None], E_idx)[..., 0]
    if trace is not None:
        trace["E_idx"], trace["h_E0"], trace["h_V0"] = E_idx, h_E, h_V
    for l in range(3):
        h_V, h_E = encoder_layer(sd, f"encoder_layers.{l}", h_V, h_E, E_idx, maskf, mask_attend, c)
        if trace is not None:
            trace[f"enc{l}_h_V"], trace[f"enc{l}_h_E"] = h_V, h_E
    h_S = sd["W_s.weight"][cg_z]
    h_S_j = _gather_nodes(h_S, E_idx)
    fixed = torch.cat([h_E, h_S_j, _gather_nodes(h_V, E_idx)], -1)      # h_EXV_encoder (:232-233)
    for l in range(3):
        h_ESV = torch.cat([h_E, h_S_j, _gather_nodes(h_V, E_idx)], -1) + fixed
        h_V = decoder_layer(sd, f"decoder_layers.{l}", h_V, h_ESV, maskf, c)
        if trace is not None:
            trace[f"dec{l}_h_V"] = h_V
    m = lin(sd, "W_out.adaLN_modulation.1", F.silu(c))[:, None, :].chunk(2, dim=-1)
    return lin(sd, "W_out.linear", _mod(_ln(h_V), m[0], m[1]))


def sample_loop(sd, z, X, cg_z, mask, noises, sched, k_neighbors=64, steps=None, keep=False):
    """GaussianDiffusion.p_sample_loop (gaussian_diffusion.py:451-547) driven through
    _WrappedModel's timestep remap (respace.py:117-129).  `noises[s]` is the N(0,1) draw used at
    respaced step s (the reference draws it with randn_like at :440).  Returns final latent
    (and the per-step samples when keep=True)."""
    T = len(sched["betas"])
    graph = edge_embedding(sd, X, mask.to(torch.int32), k_neighbors)
    graph = (graph[0], graph[3])
    x, hist = z, []
    order = list(range(T))[::-1] if steps is None else steps
    for s in order:
        t = torch.full((x.shape[0],), int(sched["timestep_map"][s]), dtype=torch.int64)
        out = denoiser_forward(sd, x, t, X, cg_z, mask, k_neighbors, graph=graph)
        x = p_sample_update(x, out, s, noises[s], sched)
        if keep:
            hist.append(x)
    return (x, hist) if keep else x


# --------------------------------------------------------------------------------------
# VQ codebook lookup (eval)
# --------------------------------------------------------------------------------------
def vq_nearest(x, codebook):
    """Nearest-code search of vector_quantize_pytorch==1.21.7 EuclideanCodebook.forward (eval):
    dist = -cdist(x, e), cdist = sqrt(clamp(|x|^2 + |e|^2 - 2 x.e, min=0)), argmax (first index
    on ties).  In-repo equivalent: utils/vq_module.py:61-66.  fp32, fixed association order
    ((x0^2+x1^2)+x2^2, ((x0 e0 + x1 e1) + x2 e2) * -2, (|x|^2+|e|^2) + xy), no FMA -- the CUDA
    kernel reproduces these bits.  x [N,3], codebook [M,3] -> idx [N] int64."""
    x = x.to(torch.float32)
    e = codebook.to(torch.float32)
    xs, es = x * x, e * e
    x2 = (xs[:, 0] + xs[:, 1]) + xs[:, 2]
    e2 = (es[:, 0] + es[:, 1]) + es[:, 2]
    out = torch.empty(x.shape[0], dtype=torch.int64)
    for s in range(0, x.shape[0], 8192):
        xc = x[s:s + 8192]
        xy = ((xc[:, None, 0] * e[None, :, 0] + xc[:, None, 1] * e[None, :, 1]) + xc[:, None, 2] * e[None, :, 2]) * -2.0
        d = sqrt_rn(((x2[s:s + 8192, None] + e2[None, :]) + xy).clamp(min=0))
        out[s:s + 8192] = torch.argmax(-d, dim=-1)
    return out


def vq_eval(latent, codebook, mask=None):
    """VectorQuantize.forward eval path as called at models/vae_model.py:740,835:
    returns (quantized, indices, loss).  Masked-out positions return the un-quantised input
    and index -1.  latent [B,L,3]; codebook [M,3] (buffer `_codebook.embed[0]`)."""
    B, L, C = latent.shape
    idx = vq_nearest(latent.reshape(-1, C), codebook).reshape(B, L)
    q = codebook[idx]
    if mask is not None:
        q = torch.where(mask[..., None], q, latent)
        idx = torch.where(mask, idx, torch.full_like(idx, -1))
    return q, idx, torch.zeros(1)


class VectorQuantizeEval(torch.nn.Module):
    """Stand-in installed as `vector_quantize_pytorch.VectorQuantize` by oracle/ref_stubs.py so the
    reference's utils/vq_module.py:106-111 and models/vae_model.py:835 run unmodified (eval only).
    Buffer name follows the third-party module (`_codebook.embed` [1, M, dim])."""

    class _CB(torch.nn.Module):
        def __init__(self, size, dim):
            super().__init__()
            embed = torch.empty(1, size, dim)
            torch.nn.init.kaiming_uniform_(embed)
            self.register_buffer("embed", embed)

    def __init__(self, dim, codebook_size, decay=0.99, commitment_weight=0.25, **_):
        super().__init__()
        self._codebook = self._CB(codebook_size, dim)

    def forward(self, x, mask=None):
        return vq_eval(x, self._codebook.embed[0], mask)


# --------------------------------------------------------------------------------------
# IC decoder (models/vae_model.py:318-503, models/gcn_nn.py:222-381)
# --------------------------------------------------------------------------------------
def _swish(x):
    return x * torch.sigmoid(x)


def _seq(sd, p, x, first=1, second=3):
    """Sequential(swish, Linear, swish, Linear) with Linear children at indices 1 and 3."""
    return lin(sd, f"{p}.{second}", _swish(lin(sd, f"{p}.{first}", _swish(x))))


def directed_edges(nbr):
    """gcn_nn.py:54-64."""
    a, b = nbr[:, 0], nbr[:, 1]
    if bool((a > b).any()) and bool((b > a).any()):
        return nbr
    return torch.cat([nbr, nbr.flip(1)], dim=0)


def ic_decoder(sd, cg_z, cg_xyz, nbr_list, S36, angle_variant=False, cutoff=21.0, p="equivaraintconv"):
    """IC_Decoder.forward (N6; vae_model.py:467-503) / IC_Decoder_angle.forward (K3,K4; :375-412).
    cg_z [N] int64, cg_xyz [N,3], nbr_list [E,2], S36 [N,36] -> ic_recon [N,13,3]."""
    nbr = directed_edges(nbr_list)
    i, j = nbr[:, 0], nbr[:, 1]
    r = cg_xyz[j] - cg_xyz[i]
    dist = ((r * r + 1e-8).sum(-1)) ** 0.5
    n = torch.arange(1, 16, dtype=torch.float32)
    coef = n * np.pi / cutoff
    d = dist[:, None]
    basis = torch.where(d >= cutoff, torch.zeros(()), torch.where(d == 0, coef, torch.sin(coef * d)) / torch.where(d == 0, torch.ones(()), d))
    env = torch.where(dist >= cutoff, torch.zeros(()), 0.5 * (torch.cos(np.pi * dist / cutoff) + 1))
    S = torch.cat([S36, sd[f"{p}.res_embed.weight"][cg_z]], dim=-1)
    N = S.shape[0]
    for b in range(4):
        mb = f"{p}.message_blocks.{b}"
        phi = lin(sd, f"{mb}.inv_dense.1", _swish(lin(sd, f"{mb}.inv_dense.0", S)))[j]
        w = lin(sd, f"{mb}.dist_embed.block.1", basis) * env[:, None]
        v = torch.zeros(N, S.shape[1]).index_add_(0, i, phi * w)
        S = S + _seq(sd, f"{p}.dense_blocks.{b}", v)
    bb_dist = sd[f"{p}.backbone_dist.weight"][cg_z]
    sc_dist = sd[f"{p}.sidechain_dist.weight"][cg_z]
    bb_angle = _seq(sd, f"{p}.backbone_angle", S)
    bb_tors = _seq(sd, f"{p}.backbone_torsion", torch.cat([S, bb_angle], -1))
    if angle_variant:
        sc_angle = _seq(sd, f"{p}.sidechain_angle", S)
        T = torch.cat([S, sc_angle], -1)
    else:
        sc_angle = sd[f"{p}.sidechain_angle.weight"][cg_z]
        T = S
    for b in range(4):
        T = T + _seq(sd, f"{p}.sidechain_torsion_blocks.{b}", T)
    sc_tors = _seq(sd, f"{p}.final_torsion", T)
    bb = torch.stack([bb_dist, bb_angle, bb_tors], -1)
    sc = torch.stack([sc_dist, sc_angle, sc_tors], -1)
    return torch.cat([bb, sc], dim=-2)


def latent_decode(sd, latent, mask, cg_z, cg_xyz, nbr_list, num_CGs, angle_variant=False):
    """VAE.latent_decode (vae_model.py:830-839) + VAE.decoder (:759-764): quantise, un-pad
    (gcn_nn.py:45-52), map_out 3->36, IC decoder.  Returns (ic_recon [sumL,13,3], idx [B,L])."""
    q, idx, _ = vq_eval(latent, sd["quantize._codebook.embed"][0], mask)
    flat = torch.cat([q[b, : int(n)] for b, n in enumerate(num_CGs)], dim=0)
    S36 = lin(sd, "map_out", flat)
    return ic_decoder(sd, cg_z, cg_xyz, nbr_list, S36, angle_variant), idx


# --------------------------------------------------------------------------------------
# Internal coordinates -> Cartesian (utils/utils_ic.py:197-268)
# --------------------------------------------------------------------------------------
def _rodrigues(axis, angle, v):
    """Rotate v by `angle` about `axis` with the Euler-Rodrigues matrix of utils_ic.py:197-210."""
    axis = axis / torch.sqrt((axis * axis).sum(-1, keepdim=True))
    a = torch.cos(angle / 2)
    s = -axis * torch.sin(angle / 2)[..., None]
    b, c, d = s[..., 0], s[..., 1], s[..., 2]
    rows = (
        (a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)),
        (2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)),
        (2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c),
    )
    return torch.stack([(r[0] * v[..., 0] + r[1] * v[..., 1]) + r[2] * v[..., 2] for r in rows], -1)


def place_atom(ic, p1, p2, p3):
    """utils_ic.py:213-239; ic [...,3] = (bond, angle, torsion)."""
    a, b = p2 - p1, p2 - p3
    a = torch.where(a == 0.0, a + 1e-8, a)
    b = torch.where(b == 0.0, b + 1e-8, b)
    d = torch.abs(ic[..., 0:1]) * a / torch.sqrt((a * a).sum(-1, keepdim=True))
    d = _rodrigues(torch.linalg.cross(a, b), ic[..., 1], d)
    d = _rodrigues(a, ic[..., 2], d)
    return p1 + d


def ic_to_slots(ca_full, ic_recon, atom_orders):
    """utils_ic.py:242-266 up to (not including) the compaction: returns the 14 slots per residue
    [B,L,14,3] in the order O, N, C, CA, side chain 0..9.  ca_full [B,L+2,3], ic_recon [B,L,13,3],
    atom_orders [10,L,3] int64."""
    ca, prv, nxt = ca_full[:, 1:-1], ca_full[:, :-2], ca_full[:, 2:]
    N = place_atom(ic_recon[:, :, 0], ca, prv, nxt)
    C = place_atom(ic_recon[:, :, 1], ca, nxt, prv)
    O = place_atom(ic_recon[:, :, 2], C, ca, N)
    slots = [O, N, C, ca]
    B, L = ca.shape[:2]
    ar = torch.arange(L)
    for s in range(10):
        cur = torch.stack(slots, dim=2)                              # [B,L,4+s,3]
        p1 = cur[:, ar, atom_orders[s, :, 2]]
        p2 = cur[:, ar, atom_orders[s, :, 1]]
        p3 = cur[:, ar, atom_orders[s, :, 0]]
        slots.append(place_atom(ic_recon[:, :, 3 + s], p1, p2, p3))
    return torch.stack(slots, dim=2)


def ic_to_xyz(CG_nxyz, ic_recon, info):
    """utils_ic.py:242-268 (full): CG_nxyz [B,L+2,4], ic_recon [B,L,13,3],
    info = (permute, atom_idx, atom_orders) -> [B,Na,3]."""
    permute, atom_idx, atom_orders = info
    slots = ic_to_slots(CG_nxyz[:, :, 1:], ic_recon, atom_orders)
    flat = slots.reshape(slots.shape[0], -1, 3)
    return flat[:, atom_idx][:, permute]

# --------------------------------------------------------------------------------------
# Evaluation step after the path
# --------------------------------------------------------------------------------------
COV_RADIUS = {1: 0.23, 6: 0.68, 7: 0.68, 8: 0.68, 15: 0.75, 16: 1.02, 34: 1.22}     # COVCUTOFFTABLE, utils/protein_module.py:128-


def bond_graph(xyz, z, scale=1.3):
    """get_bond_graphs (utils/protein_module.py:279-288): dense boolean adjacency, dist < (r_i + r_j) * scale, zero diagonal.
    fp32 like the reference (torch.Tensor(positions)); the square root is the correctly rounded one (see sqrt_rn)."""
    xyz = xyz.to(torch.float32)
    r = torch.tensor([COV_RADIUS[int(v)] for v in z.tolist()], dtype=torch.float32)
    dist = sqrt_rn((xyz[:, None, :] - xyz[None, :, :]).pow(2).sum(-1))
    bond = dist < (r[None, :] + r[:, None]) * scale
    bond.fill_diagonal_(False)
    return bond


def sample_quality_stats(xyz_ref, xyz_gen, z, num_atoms, scale=1.3):
    """What eval_sample_qualities (utils/protein_module.py:335-364) derives its outputs from, per structure:
    counts [n, 6] = {#entries where the graphs differ, #ref bonds, #gen bonds} for all atoms, then for heavy atoms only
    (dropH, :266-277), and sums [n, 4] = {sum of squared deviations, Na, the same over heavy atoms, #heavy} (compute_rmsd, :319-333)."""
    counts, sums, o = [], [], 0
    for na in [int(v) for v in num_atoms]:
        r, g, zz = xyz_ref[o:o + na], xyz_gen[o:o + na], z[o:o + na]
        heavy = zz != 1
        row = []
        for sel in (torch.ones_like(heavy), heavy):
            br, bg = bond_graph(r[sel], zz[sel], scale), bond_graph(g[sel], zz[sel], scale)
            row += [int((br != bg).sum()), int(br.sum()), int(bg.sum())]
        d2 = (g.double() - r.double()).pow(2).sum(-1)
        counts.append(row)
        sums.append([float(d2.sum()), float(na), float(d2[heavy].sum()), float(heavy.sum())])
        o += na
    return torch.tensor(counts, dtype=torch.int64), torch.tensor(sums, dtype=torch.float64)


def superposed_rmsd(a, b):
    """md.rmsd(Trajectory(a), Trajectory(b)) as test.py:37-79 uses it: minimum RMSD of two [Na, 3] structures under rigid
    superposition.  mdtraj (absent here: parity unpinned against it) computes it with Theobald's QCP in float32; this is the
    float64 Kabsch form of the same quantity: rmsd^2 = (|a_c|^2 + |b_c|^2 - 2 sum_i sigma_i d_i) / N with d = (1, 1, sign det)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    ac, bc = a - a.mean(0), b - b.mean(0)
    u, sig, vt = np.linalg.svd(ac.T @ bc)
    d = np.ones(3)
    d[2] = np.sign(np.linalg.det(u @ vt))
    return float(np.sqrt(max(0.0, ((ac ** 2).sum() + (bc ** 2).sum() - 2.0 * (sig * d).sum()) / a.shape[0])))


def compute_div(gen_structures, ref_structure):
    """compute_div (test.py:82-96) on numpy arrays: 1 - mean rmsd(gen, mean gen) / mean rmsd(gen, ref)."""
    gen = np.asarray(gen_structures, dtype=np.float64)
    ref = np.asarray(ref_structure, dtype=np.float64)
    mean_gen = gen.mean(0)
    r_ref = np.mean([superposed_rmsd(gen[i][p], ref[p]) for i in range(gen.shape[0]) for p in range(gen.shape[1])])
    r_gen = np.mean([superposed_rmsd(gen[i][p], mean_gen[p]) for i in range(gen.shape[0]) for p in range(gen.shape[1])])
    return 1.0 - r_gen / r_ref, r_ref, r_gen


# ---------------------------------------------------------------------------------------------- pair-list evaluation losses
EVAL_EPS = 1e-7                                                     # test.py:27


def _pd(xyz, pairs):
    return ((xyz[pairs[:, 0]] - xyz[pairs[:, 1]]).pow(2).sum(-1) + EVAL_EPS).sqrt()


def inter_result(interaction_list, pi_pi_list, xyz_recon):
    """test.py:97-116."""
    n_inter, n_pi = interaction_list.shape[0], pi_pi_list.shape[0]
    tot = n_inter + n_pi
    loss_inter = torch.tensor(0.0)
    loss_pi = torch.tensor(0.0)
    if n_inter > 0:
        loss_inter = torch.clamp_min(_pd(xyz_recon, interaction_list) - 4.0, 0.0).mean() * (n_inter / tot)
    if n_pi > 0:
        c0 = (xyz_recon[pi_pi_list[:, 0]] + xyz_recon[pi_pi_list[:, 1]]) / 2
        c1 = (xyz_recon[pi_pi_list[:, 2]] + xyz_recon[pi_pi_list[:, 3]]) / 2
        loss_pi = torch.clamp_min(((c0 - c1).pow(2).sum(-1) + EVAL_EPS).sqrt() - 6.0, 0.0).mean()
        loss_inter = loss_inter + loss_pi * (n_pi / tot)
    return loss_inter, loss_pi


def clash_pairs(edge_list, nbr_list):
    """test.py:120-123: rows occurring exactly once in the concatenation."""
    u, c = torch.cat((edge_list, nbr_list)).unique(dim=0, return_counts=True)
    return u[c == 1]


def clash_result(edge_list, nbr_list, xyz_recon, bb_NO_list):
    """test.py:118-139."""
    d = _pd(xyz_recon, clash_pairs(edge_list, nbr_list))
    loss = (d < 1.2).sum().float() / d.numel() if d.numel() > 0 else torch.tensor(0.0)
    b = _pd(xyz_recon, bb_NO_list)
    return loss + ((b < 1.2).sum().float() / b.numel() if b.numel() > 0 else torch.tensor(0.0))


def ged_result(xyz_recon, xyz, edge_list):
    """test.py:141-146."""
    return (_pd(xyz_recon, edge_list) - _pd(xyz, edge_list)).pow(2).mean()
