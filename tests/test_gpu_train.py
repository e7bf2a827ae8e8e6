"""GPU: the train_latent step -- CUDA forward / backward of the denoiser composed from the cb2t_* operators
(include/codlad_b200_train.h) against torch.autograd on the CPU oracle (oracle/restate.py, itself pinned against the unmodified
reference), the fused AdamW + EMA kernel against torch.optim.AdamW + the reference's update_ema, and a few whole steps."""
import numpy as np
import pytest
import torch

from codlad_b200 import synthetic, weights
from tests import parity_utils as P

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    from oracle import restate
    return restate


def _ragged_batch(lengths, seed):
    prots = [synthetic.make_protein(n, 1, seed=seed + i) for i, n in enumerate(lengths)]
    batch = synthetic.collate_many(prots)
    Lmax = max(lengths)
    X = torch.zeros(len(lengths), Lmax, 3)
    z = torch.zeros(len(lengths), Lmax, dtype=torch.int64)
    for b, (p, n) in enumerate(zip(prots, lengths)):
        X[b, :n] = p.ca_full[0, 1:-1]
        z[b, :n] = p.restype_full[1:-1]
    mask = torch.arange(Lmax)[None, :] < torch.tensor(lengths)[:, None]
    return batch, X, z, mask


@pytest.mark.parametrize("M,N,K,a_kc,b_kc", [(300, 128, 128, 1, 1), (257, 512, 128, 1, 1), (300, 128, 384, 1, 0), (128, 128, 9001, 0, 0),
                                             (128, 151, 5000, 0, 0), (77, 6, 128, 1, 1), (128, 3, 4500, 0, 0), (65, 128, 16, 1, 1)])
def test_gemm_against_float64(M, N, K, a_kc, b_kc):
    """cb2t_gemm in its three roles (forward / dgrad / wgrad incl. the split reduction and odd leading dimensions) vs float64 on the CPU."""
    from codlad_b200 import _native as Nn
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((M, K) if a_kc else (K, M), generator=g)
    B = torch.randn((N, K) if b_kc else (K, N), generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = (A.double() if a_kc else A.double().t()) @ (B.double().t() if b_kc else B.double())
    Ad, Bd, Cd = A.cuda(), B.cuda(), C0.cuda().clone()
    Nn.check(Nn.lib().cb2t_set_gemm_mode(0))
    for acc in (0, 1):
        Cd.copy_(C0)
        Nn.check(Nn.lib().cb2t_gemm(Ad.data_ptr(), Bd.data_ptr(), Cd.data_ptr(), M, N, K, A.shape[1], B.shape[1], N, a_kc, b_kc, acc, Nn.stream_ptr()))
        want = ref + (C0.double() if acc else 0)
        err = float((Cd.cpu().double() - want).abs().max() / want.abs().max())
        assert err < 2e-6 * max(1.0, K / 1000) ** 0.5, (acc, err)


@pytest.mark.parametrize("M,N,K,a_kc,b_kc", [(4096, 128, 128, 1, 1), (5000, 512, 128, 1, 1), (3001, 128, 512, 1, 1), (2000, 1152, 128, 1, 1),
                                             (4100, 128, 152, 1, 1), (128, 128, 16384, 0, 0), (512, 128, 70001, 0, 0), (128, 512, 9000, 0, 0)])
def test_gemm_tf32_tensor_cores_against_float64(M, N, K, a_kc, b_kc):
    """cb2t_gemm in TF32 mode (tcgen05.mma kind::tf32): the K-major form (forward / data gradient) with M, N, K tails, and the MN-major form
    (weight gradient: 32-byte-atom swizzle, split reduction).  Bar 2e-3 of the largest output (TF32 rounds operands to 10 mantissa bits;
    measured 7-9e-4)."""
    from codlad_b200 import _native as Nn
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((M, K) if a_kc else (K, M), generator=g)
    B = torch.randn((N, K) if b_kc else (K, N), generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = (A.double() if a_kc else A.double().t()) @ (B.double().t() if b_kc else B.double())
    Ad, Bd, Cd = A.cuda(), B.cuda(), C0.cuda().clone()
    try:
        Nn.check(Nn.lib().cb2t_set_gemm_mode(1))
        for acc in (0, 1):
            Cd.copy_(C0)
            Nn.check(Nn.lib().cb2t_gemm(Ad.data_ptr(), Bd.data_ptr(), Cd.data_ptr(), M, N, K, A.shape[1], B.shape[1], N, a_kc, b_kc, acc, Nn.stream_ptr()))
            want = ref + (C0.double() if acc else 0)
            err = float((Cd.cpu().double() - want).abs().max() / want.abs().max())
            assert 1e-5 < err < 2e-3, (acc, err)          # > 1e-5: the tensor-core path really ran (fp32 SIMT would be ~5e-7)
    finally:
        Nn.check(Nn.lib().cb2t_set_gemm_mode(0))


@pytest.mark.parametrize("mode,M,N,K,act", [(1, 4096, 128, 128, True), (1, 5001, 512, 128, True), (1, 3000, 128, 512, False), (1, 2048, 64, 96, True),
                                            (0, 700, 128, 128, True), (0, 300, 128, 128, False), (1, 500, 128, 128, True)])
def test_fused_linear_bias_gelu_against_float64(mode, M, N, K, act):
    """cb2t_linear_bias_gelu_fwd: Z = X W^T + b and Y = GELU(Z).  In TF32 mode (large M) bias and GELU run in the tensor-core GEMM's epilogue
    (two stores per accumulator block, M tails clipped by the TMA stores); small M or fp32 mode take the GEMM + elementwise pair.  Both outputs
    against float64; Y must be GELU of the Z that was stored (bit-for-bit the same erf formula)."""
    from codlad_b200 import _native as Nn
    g = torch.Generator().manual_seed(M * 7 + N + K)
    X, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    Zref = X.double() @ W.double().t() + b.double()
    Xd, Wd, bd = X.cuda(), W.cuda(), b.cuda()
    Z = torch.full((M, N), float("nan"), device="cuda")
    Y = torch.full((M, N), float("nan"), device="cuda") if act else None
    try:
        Nn.check(Nn.lib().cb2t_set_gemm_mode(mode))
        Nn.check(Nn.lib().cb2t_linear_bias_gelu_fwd(Xd.data_ptr(), Wd.data_ptr(), bd.data_ptr(), Z.data_ptr(), Y.data_ptr() if act else None, M, N, K, K, K, N,
                                                    Nn.stream_ptr()))
    finally:
        Nn.check(Nn.lib().cb2t_set_gemm_mode(0))
    bar = 2e-3 if mode == 1 and M >= 1024 else 3e-6
    err = float((Z.cpu().double() - Zref).abs().max() / Zref.abs().max())
    assert err < bar, err
    if mode == 1 and M >= 1024:
        assert err > 1e-5                                    # the tensor-core path really ran
    if act:
        want = torch.nn.functional.gelu(Z.cpu().double())
        assert float((Y.cpu().double() - want).abs().max()) < 2e-6


def test_gelu_backward_with_bias_gradient():
    """cb2t_gelu_bwd_colsum = cb2t_gelu_bwd followed by cb2t_colsum, in one pass: identical dZ, bias gradient equal to the separate column sum
    (same partition and order: bit-identical) and to float64 within fp32 summation error; accumulates into the gradient."""
    from codlad_b200 import _native as Nn
    g = torch.Generator().manual_seed(77)
    for rows, cols in ((50000, 128), (3001, 512), (17, 128)):
        pre, dY = torch.randn(rows, cols, generator=g).cuda(), torch.randn(rows, cols, generator=g).cuda()
        d1, d2 = dY.clone(), dY.clone()
        b0 = torch.randn(cols, generator=g).cuda()
        b1, b2 = b0.clone(), b0.clone()
        lib = Nn.lib()
        Nn.check(lib.cb2t_gelu_bwd(pre.data_ptr(), d1.data_ptr(), rows * cols, d1.data_ptr(), Nn.stream_ptr()))
        Nn.check(lib.cb2t_colsum(d1.data_ptr(), rows, cols, cols, b1.data_ptr(), 1, Nn.stream_ptr()))
        Nn.check(lib.cb2t_gelu_bwd_colsum(pre.data_ptr(), d2.data_ptr(), rows, cols, d2.data_ptr(), b2.data_ptr(), 1, Nn.stream_ptr()))
        assert torch.equal(d1, d2) and torch.equal(b1, b2)
        want = b0.double().cpu() + d1.double().cpu().sum(0)
        assert float((b2.cpu().double() - want).abs().max()) < 1e-4 * max(1.0, float(want.abs().max()))


def test_denoiser_gradients_tf32_mode(R):
    """The same backward with the large GEMMs on the tensor cores in TF32 -- the arithmetic the reference trains with (train_latent.py:24-25).
    Against fp32 autograd every gradient tensor stays within 1e-2 of its norm (measured ~2e-3), the model output within 5e-3."""
    from codlad_b200 import train
    sd = weights.init_denoiser_state(7)
    lengths = [72, 66, 70]
    batch, X, z, mask = _ragged_batch(lengths, 600)
    B, L = mask.shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, L, 3, generator=g)
    t = torch.tensor([999, 412, 3])
    dout = torch.randn(B, L, 6, generator=g) * mask[..., None]
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        out_ref = R.denoiser_forward(leaves, x, t, X, z, mask, 64)
        (out_ref * dout).sum().backward()
    tr = train.DenoiserTrainer(sd, gemm="tf32")
    try:
        geom = train.Geometry(batch, 64, tr.device)
        out = tr.forward(x.cuda(), t.cuda(), geom)
        assert float((out.cpu() - out_ref.detach())[mask].abs().max()) < 5e-3
        tr.zero_grad()
        tr.backward(dout.cuda())
    finally:
        N_ = __import__("codlad_b200._native", fromlist=["x"])
        N_.check(N_.lib().cb2t_set_gemm_mode(0))
    worst = ("", 0.0)
    for k, leaf in leaves.items():
        ref = leaf.grad if leaf.grad is not None else torch.zeros_like(leaf)
        got = tr.grads[k].cpu()
        if k == "features.embeddings.linear.weight":
            ref, got = ref[:, :65], got[:, :65]
        err = float((got - ref).norm()) / (float(ref.norm()) + 1e-12)
        worst = max(worst, (k, err), key=lambda kv: kv[1])
        assert err < 1e-2, (k, err)
    print(f"tf32 mode: worst gradient error {worst[1]:.2e} ({worst[0]})")


@pytest.mark.parametrize("lengths,kn", [([40, 33], 64), ([72, 66, 70], 64), ([50], 32)])
def test_denoiser_gradients_match_autograd_on_the_oracle(R, lengths, kn):
    """Every one of the 108 parameter gradients of the WHOLE denoiser (3 encoder layers with node message, FFN and edge update, 3 decoder
    layers, featuriser projections, timestep embedder, FinalLayer), ragged batch, K < 64 and K = 64: CUDA backward vs torch.autograd through
    the oracle's forward (p = 0 dropout).  Bar: 2e-5 of each tensor's gradient norm (fp32 both sides; measured 2.4e-6)."""
    from codlad_b200 import train
    sd = weights.init_denoiser_state(7)
    batch, X, z, mask = _ragged_batch(lengths, 600)
    B, L = mask.shape
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, L, 3, generator=g)
    t = torch.tensor([999, 412, 3][:B])
    dout = torch.randn(B, L, 6, generator=g) * mask[..., None]
    # oracle + autograd (CPU fp32)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    with torch.enable_grad():
        out_ref = R.denoiser_forward(leaves, x, t, X, z, mask, kn)
        (out_ref * dout).sum().backward()
    tr = train.DenoiserTrainer(sd, k_neighbors=kn)
    geom = train.Geometry(batch, kn, tr.device)
    out = tr.forward(x.cuda(), t.cuda(), geom)
    m = mask
    assert float((out.cpu() - out_ref.detach())[m].abs().max()) < 5e-5
    tr.zero_grad()
    tr.backward(dout.cuda())
    worst = ("", 0.0)
    for k, leaf in leaves.items():
        ref = leaf.grad if leaf.grad is not None else torch.zeros_like(leaf)
        got = tr.grads[k].cpu()
        if k == "features.embeddings.linear.weight":
            ref, got = ref[:, :65], got[:, :65]          # class 65 is never produced (chain mask == 1)
        denom = float(ref.norm()) + 1e-12
        err = float((got - ref).norm()) / denom if denom > 1e-10 else float(got.norm())
        if err > worst[1]:
            worst = (k, err)
        assert err < 2e-5, (k, err, denom)
    print(f"lengths {lengths}: worst gradient error {worst[1]:.2e} ({worst[0]})")


def test_adamw_ema_matches_torch(R):
    """cb2t_sumsq + cb2t_adamw_ema == clip_grad_norm_(1.0) + torch.optim.AdamW(lr, weight_decay=0) + update_ema(decay) (train_latent.py:114,
    252-261), five steps on random gradients."""
    from codlad_b200 import train
    sd = weights.init_denoiser_state(1)
    tr = train.DenoiserTrainer(sd, lr=3e-3, ema_decay=0.99, grad_clip=1.0)
    ref_p = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items()}
    ref_ema = {k: v.clone() for k, v in sd.items()}
    opt = torch.optim.AdamW(list(ref_p.values()), lr=3e-3, weight_decay=0)
    g = torch.Generator().manual_seed(5)
    for step in range(5):
        scale = 10.0 if step % 2 == 0 else 1e-3          # with and without the clip being active
        for k, p in ref_p.items():
            gr = torch.randn(p.shape, generator=g) * scale
            p.grad = gr.clone()
            tr.grads[k].copy_(gr)
        norm = torch.nn.utils.clip_grad_norm_(list(ref_p.values()), 1.0)
        opt.step()
        for k in ref_ema:
            ref_ema[k].mul_(0.99).add_(ref_p[k].data, alpha=0.01)
        tr.step()
        assert abs(tr.grad_norm() - float(norm)) / float(norm) < 1e-5
    for k in sd:
        assert torch.allclose(tr.params[k].cpu(), ref_p[k].data, rtol=2e-5, atol=2e-7), k
        assert torch.allclose(tr.ema[k].cpu(), ref_ema[k], rtol=2e-5, atol=2e-7), k


def test_train_step_loss_and_descent(R):
    """One whole step as train_latent.py:203-261 runs it: the loss terms equal training_losses fed with the oracle's forward, and a few
    AdamW steps on a fixed batch reduce the loss; the EMA trails the parameters."""
    from codlad_b200 import train
    from codlad_b200.diffusion import create_diffusion
    sd = weights.init_denoiser_state(2)
    lengths = [48, 40]
    batch, X, z, mask = _ragged_batch(lengths, 700)
    diffusion = create_diffusion("")
    g = torch.Generator().manual_seed(3)
    x1 = torch.randn(2, 48, 3, generator=g)
    noise = torch.randn(2, 48, 3, generator=g)
    t = torch.tensor([250, 800])
    tr = train.DenoiserTrainer(sd, lr=2e-3)
    geom = train.Geometry(batch, 64, tr.device)
    terms, loss0 = tr.train_step(diffusion, x1, t, batch, noise=noise, geom=geom, do_step=False)
    ref_model = lambda x_t, tt, **kw: R.denoiser_forward(sd, x_t, tt, X, z, mask, 64)
    with torch.no_grad():
        ref = diffusion.training_losses(ref_model, x1, t, dict(mask=mask), noise=noise)
    for k in ("mse", "vb", "loss"):
        assert torch.allclose(terms[k].cpu(), ref[k], rtol=1e-3, atol=1e-5), (k, terms[k], ref[k])
    assert torch.is_tensor(loss0) and loss0.is_cuda and loss0.dim() == 0        # a device scalar like the reference's `loss`: no sync in the step
    losses = [float(loss0)]
    for _ in range(12):
        _, l = tr.train_step(diffusion, x1, t, batch, noise=noise, geom=geom)
        losses.append(float(l))
    print("losses", [round(v, 4) for v in losses])
    assert losses[-1] < 0.8 * losses[0]
    assert not torch.equal(tr.flat_ema, tr.flat_p) and float((tr.flat_ema - tr.flat_p).abs().max()) > 0
    # dropout path (the reference's p = 0.6) runs and stays finite
    _, ld = tr.train_step(diffusion, x1, t, batch, noise=noise, geom=geom, dropout_p=0.6, generator=torch.Generator(device="cuda").manual_seed(1))
    assert np.isfinite(float(ld)) and torch.isfinite(tr.flat_p).all()
