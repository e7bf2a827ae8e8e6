"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the CUDA path through the
C ABI (codlad_b200.engine -> libcodlad_b200.so) and checks it against the CPU oracle
(oracle/restate.py) on the same seeded inputs and against the committed reference goldens.

Bars (BASELINE.json north_star): k-NN indices and VQ code indices bit-exact; floating-point stages
within the tolerance written next to each assert (fp32 tier ~1e-5 relative; coordinates 1e-3 A RMSD
given identical indices; bf16 tier 5e-2 A)."""
import numpy as np
import pytest
import torch

from codlad_b200 import synthetic, weights
from tests import parity_utils as P

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


@pytest.fixture(scope="module")
def eng():
    from codlad_b200 import engine
    return engine


@pytest.fixture(scope="module")
def R():
    from oracle import restate
    return restate


@pytest.fixture(scope="module")
def denoiser(eng):
    sd = weights.init_denoiser_state(0)
    return sd, eng.DenoiserEngine(sd, 64)


def _plan_for_case(eng, den, c, precision="fp32", keep_debug=True, k_neighbors=64):
    B, L = c["mask"].shape
    d = den if k_neighbors == 64 else eng.DenoiserEngine(weights.init_denoiser_state(0), k_neighbors)
    plan = eng.Plan(d, B, B, L, precision, keep_debug)
    plan._keep = d
    plan.set_frames(c["X"], c["mask"].sum(1).int(), c["cg_z"].int(), torch.arange(B, dtype=torch.int32))
    return plan


# --------------------------------------------------------------------------------------- k-NN
@pytest.mark.parametrize("L,F,K,seed", [(64, 2, 64, 11), (300, 3, 64, 12), (2000, 1, 48, 13), (33, 2, 33, 14), (500, 4, 64, 15)])
def test_knn_bit_exact(eng, R, L, F, K, seed):
    X = synthetic.ca_trace(F, L, seed)
    D_ref, I_ref = R.knn_graph(X, torch.ones(F, L), K)
    D, I = eng.knn_topk(X.cuda(), None, K)
    assert torch.equal(D.cpu(), D_ref), "distances must be bit-exact"
    assert torch.equal(I.cpu().long(), I_ref), "indices must be bit-exact (ties broken by lowest index)"


def test_knn_ragged_padding(eng, R):
    L = 96
    lengths = torch.tensor([96, 70, 40])
    X = synthetic.ca_trace(3, L, 21)
    mask = torch.arange(L)[None, :] < lengths[:, None]
    X = X * mask[..., None]
    D_ref, I_ref = R.knn_graph(X, mask.float(), 64)
    D, I = eng.knn_topk(X.cuda(), lengths, 64)
    assert torch.equal(D.cpu(), D_ref) and torch.equal(I.cpu().long(), I_ref)


def test_knn_against_reference_golden(eng):
    g = P.golden("denoiser_L100_K48")
    c = P.denoiser_case(g["meta"])
    D, I = eng.knn_topk(c["X"].cuda(), None, 48)
    ulp = np.abs(D.cpu().numpy().view(np.int32).astype(np.int64) - g["D_nb"].view(np.int32).astype(np.int64))
    assert ulp.max() <= 1 and (ulp != 0).mean() < 0.02     # golden = torch-CPU sqrt (1 ulp low on ~0.6 % of inputs)
    assert P.knn_tie_aware_equal(I.cpu().numpy().reshape(-1, 48), g["E_idx"].reshape(-1, 48), g["D_nb"].reshape(-1, 48))


# --------------------------------------------------------------------------------------- features + denoiser (fp32 tier)
CASES = [("denoiser_L64_B1", None), ("denoiser_L70_B2", None), ("denoiser_L100_K48", None),
         ("denoiser_L40_short", None), ("denoiser_ragged", [80, 70])]


@pytest.mark.parametrize("name,lengths", CASES)
def test_edge_features_fp32(eng, R, denoiser, name, lengths):
    sd, den = denoiser
    g = P.golden(name)
    c = P.denoiser_case(g["meta"], lengths)
    plan = _plan_for_case(eng, den, c, k_neighbors=c["k_neighbors"])
    E_idx, D_nb, E, hE0 = R.edge_embedding(sd, c["X"], c["mask"].int(), c["k_neighbors"])
    m = c["mask"]
    assert torch.equal(plan.buffer("nbr_idx").cpu().long()[m], E_idx[m])
    E_gpu, h_gpu = plan.buffer("E").cpu(), plan.buffer("hE0").cpu()
    # the quaternion's sqrt(|~0|) terms are ill-conditioned in the reference formula itself (~3e-4)
    assert (E_gpu - E)[m].abs().max() < 2e-3
    assert (E_gpu - E)[m].abs().mean() < 2e-6
    assert (h_gpu - hE0)[m].abs().mean() < 2e-6


@pytest.mark.parametrize("name,lengths", CASES)
def test_denoiser_forward_fp32(eng, R, denoiser, name, lengths):
    sd, den = denoiser
    g = P.golden(name)
    c = P.denoiser_case(g["meta"], lengths)
    plan = _plan_for_case(eng, den, c, k_neighbors=c["k_neighbors"])
    out = plan.forward(c["x"].cuda(), c["t"].float().cuda()).cpu()
    ref = R.denoiser_forward(sd, c["x"], c["t"], c["X"], c["cg_z"], c["mask"], c["k_neighbors"])
    m = c["mask"]
    if lengths is not None:
        K = plan.K
        m = m & torch.tensor([n >= K for n in lengths])[:, None]
    err = (out - ref)[m].abs().max().item()
    assert err < 5e-5, f"fp32 tier vs oracle: {err}"
    gerr = np.abs(out.numpy() - g["out"])[m.numpy()].max()
    assert gerr < 5e-5, f"fp32 tier vs reference golden: {gerr}"


def test_ensemble_members_share_frame(eng, R, denoiser):
    sd, den = denoiser
    prot = synthetic.make_protein(48, 1, seed=77)
    X = prot.ca_full[:, 1:-1].contiguous()
    z = prot.restype_full[1:-1][None]
    NB = 3
    plan = eng.Plan(den, 1, NB, 48, "fp32")
    plan.set_frames(X, torch.tensor([48]), z.int(), torch.zeros(NB, dtype=torch.int32))
    x = synthetic.latent_noise((NB, 48, 3), 5)
    t = torch.tensor([3.0, 500.0, 999.0])
    out = plan.forward(x.cuda(), t.cuda()).cpu()
    ref = R.denoiser_forward(sd, x, t.long(), X.expand(NB, -1, -1), z.expand(NB, -1), torch.ones(NB, 48, dtype=torch.bool))
    assert (out - ref).abs().max() < 5e-5


@pytest.mark.parametrize("L,NB", [(2, 1), (3, 2), (8, 1), (17, 3), (31, 2)])
def test_tiny_proteins_both_tiers_vs_oracle(eng, R, denoiser, L, NB):
    """The smallest inputs the reference accepts (k = min(K, L) neighbours, protein_mpnn_utils.py:454; a tile holds more nodes than the
    frame has): both tiers against the oracle forward, and k-NN exact."""
    sd, den = denoiser
    prot = synthetic.make_protein(L, 1, seed=900 + L)
    X = prot.ca_full[:, 1:-1].contiguous()
    z = prot.restype_full[1:-1][None]
    x = synthetic.latent_noise((NB, L, 3), 19)
    t = torch.linspace(0, 999, NB).round()
    ref = R.denoiser_forward(sd, x, t.long(), X.expand(NB, -1, -1), z.expand(NB, -1), torch.ones(NB, L, dtype=torch.bool))
    D_ref, I_ref = R.knn_graph(X, torch.ones(1, L), min(64, L))
    for prec, bar in (("fp32", 5e-5), ("f16", 8e-3)):
        plan = eng.Plan(den, 1, NB, L, prec)
        plan.set_frames(X, torch.tensor([L]), z.int(), torch.zeros(NB, dtype=torch.int32))
        assert torch.equal(plan.buffer("nbr_idx").cpu().long().reshape(1, L, -1), I_ref)
        out = plan.forward(x.cuda(), t.cuda()).cpu()
        assert torch.isfinite(out).all()
        assert (out - ref).abs().max() < bar, (prec, float((out - ref).abs().max()))


def test_empty_and_malformed_frame_sets_are_rejected():
    """No frames / zero-length frames / mismatched topology lists raise before any kernel runs (the reference fails inside torch ops on
    such batches; here the C ABI must never see them)."""
    from codlad_b200 import sampler
    prot = synthetic.make_protein(12, 1, seed=5)
    batch = synthetic.collate(prot)
    empty = {k: v[:0] for k, v in batch.items()}
    with pytest.raises((ValueError, RuntimeError)):
        sampler.frames_from_batch(empty, [], 1)
    with pytest.raises((ValueError, RuntimeError)):
        sampler.frames_from_batch(batch, [], 1)
    bad = dict(batch)
    bad["num_CGs"] = torch.tensor([0])
    with pytest.raises((ValueError, RuntimeError)):
        sampler.frames_from_batch(bad, [prot.info], 1)
    with pytest.raises((ValueError, RuntimeError)):
        sampler.frames_from_batch(batch, [prot.info], 0)


# --------------------------------------------------------------------------------------- tcgen05 (fp16) tier
@pytest.mark.parametrize("name,lengths", CASES)
def test_denoiser_forward_f16_tier(eng, R, denoiser, name, lengths):
    """fp16 operands / fp32 accumulation on the tensor cores, tanh-form GELU: tolerance 3e-3 relative (L2) and
    1e-2 max-abs on O(1) outputs against the fp32 CPU oracle (measured ~3e-4 / ~1e-3)."""
    sd, den = denoiser
    g = P.golden(name)
    c = P.denoiser_case(g["meta"], lengths)
    plan = _plan_for_case(eng, den, c, precision="f16", keep_debug=False, k_neighbors=c["k_neighbors"])
    out = plan.forward(c["x"].cuda(), c["t"].float().cuda()).cpu()
    ref = R.denoiser_forward(sd, c["x"], c["t"], c["X"], c["cg_z"], c["mask"], c["k_neighbors"])
    m = c["mask"]
    if lengths is not None:
        m = m & torch.tensor([n >= plan.K for n in lengths])[:, None]
    assert P.rel_err(out[m], ref[m]) < 3e-3
    assert (out - ref)[m].abs().max() < 1e-2
    assert P.rel_err(out[m], torch.from_numpy(g["out"])[m]) < 3e-3


@pytest.mark.parametrize("L,NB,kn", [(65, 3, 64), (96, 2, 48), (33, 2, 64), (130, 5, 32)])
def test_f16_tier_matches_fp32_tier_on_odd_geometries(eng, denoiser, L, NB, kn):
    """Tile tails (L not a multiple of the nodes per tile), K < 64, K not a multiple of 8, several members per frame."""
    sd, den = denoiser
    d = den if kn == 64 else eng.DenoiserEngine(sd, kn)
    prot = synthetic.make_protein(L, 1, seed=300 + L)
    X = prot.ca_full[:, 1:-1].contiguous()
    z = prot.restype_full[1:-1][None].int()
    x = synthetic.latent_noise((NB, L, 3), 9).cuda()
    t = torch.linspace(1, 999, NB).cuda()
    outs = {}
    for prec in ("fp32", "f16"):
        plan = eng.Plan(d, 1, NB, L, prec)
        plan.set_frames(X, torch.tensor([L]), z, torch.zeros(NB, dtype=torch.int32))
        outs[prec] = plan.forward(x, t).cpu()
    assert P.rel_err(outs["f16"], outs["fp32"]) < 3e-3


def test_sampler_100_steps_f16_tier(eng, denoiser):
    """100 reverse steps on the tensor-core tier: final latent within 1e-2 relative of the reference golden
    (measured ~1e-3; the fp32 tier is at ~1e-5), graph replay bit-identical to eager launches."""
    from codlad_b200.diffusion import create_diffusion
    sd, den = denoiser
    g = P.golden("sampler_L64_100")
    prot, z0, noises, steps = _sampler_inputs(g)
    diff = create_diffusion(str(steps))
    plan = eng.Plan(den, 1, 1, prot.L, "f16")
    plan.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([prot.L]), prot.restype_full[1:-1][None].int(),
                    torch.zeros(1, dtype=torch.int32))
    plan.set_schedule(diff.timestep_map, diff.coef_table())
    nz = noises.cuda().contiguous()
    x = z0.cuda().clone()
    plan.sample(x, nz, use_graph=False)
    err = P.rel_err(x.cpu(), g["sample_0"])
    print(f"f16 tier, 100 steps: latent rel err {err:.3e}")
    assert err < 1.5e-3                                 # measured 4.6e-4
    xg = z0.cuda().clone()
    plan.sample(xg, nz, use_graph=True)
    assert torch.equal(xg, x)


def test_full_path_f16_tier_stagewise(eng, R):
    """configs[1]-shaped slice (1 frame x 4 members x 120 residues, 100 steps) on the fp16 tier, graded stage-wise
    (SURVEY.md 'hard parts'): (ii) final latents vs the fp32 tier, (iii) VQ indices exact GIVEN the same latents,
    (iv) coordinates within 5e-2 A RMSD given identical indices; the end-to-end code flip rate is reported."""
    from codlad_b200 import sampler
    dsd, vsd = weights.init_denoiser_state(0), weights.init_vae_decode_state(0)
    prot = synthetic.make_protein(120, 1, seed=1010)
    batch = synthetic.collate(prot)
    ENS = 4
    fs = sampler.frames_from_batch(batch, prot.info, ENS)
    z0 = synthetic.latent_noise((ENS, 120, 3), 2010)
    noises = synthetic.latent_noise((100, ENS, 120, 3), 3010)
    res = {}
    for prec in ("fp32", "f16"):
        bm = sampler.Backmapper(dsd, vsd, "N6", precision=prec)
        plan = bm.upload(fs)
        out = bm.sample(plan, fs, z0, noises)
        res[prec] = {k: v.cpu().clone() for k, v in out.items()}
    lat_err = P.rel_err(res["f16"]["latent"], res["fp32"]["latent"])
    flips = float((res["f16"]["idx"] != res["fp32"]["idx"]).float().mean())
    print(f"f16 vs fp32 tier: latent rel err {lat_err:.3e}, end-to-end VQ code flip rate {flips:.4f}")
    assert lat_err < 1e-3                               # measured 2.8e-4
    assert flips <= 0.005                               # measured 0 (at most 2 of 480 codes)
    # (iii)/(iv): decode the f16-tier latent with the CPU oracle: indices exact, coordinates within tolerance
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    mask = torch.ones(ENS, 120, dtype=torch.bool)
    rep = lambda t: torch.cat([t] * ENS, 0)
    nbr = torch.cat([batch["CG_nbr_list"] + e * 120 for e in range(ENS)], 0)
    ic_ref, idx_ref = R.latent_decode(vsd, res["f16"]["latent"] * std + mean, mask, rep(batch["CG_nxyz"][:, 0].long()),
                                      rep(batch["CG_nxyz"][:, 1:]), nbr, torch.full((ENS,), 120), False)
    assert torch.equal(res["f16"]["idx"].long(), idx_ref)
    og = batch["OG_CG_nxyz"].reshape(1, 122, 4).expand(ENS, -1, -1)
    xyz_ref = R.ic_to_xyz(og, ic_ref.reshape(ENS, 120, 13, 3), prot.info)
    assert P.rmsd(res["f16"]["xyz"].reshape(ENS, -1, 3), xyz_ref) < 5e-2


# --------------------------------------------------------------------------------------- sampler
def _sampler_inputs(g):
    L, prot_seed, z_seed, noise_seed, steps = (int(v) for v in g["meta"])
    prot = synthetic.make_protein(L, 1, seed=prot_seed)
    z0 = synthetic.latent_noise((1, L, 3), z_seed)
    noises = synthetic.latent_noise((steps, 1, L, 3), noise_seed)
    return prot, z0, noises, steps


def test_sampler_100_steps_fp32_vs_reference_golden(eng, denoiser):
    from codlad_b200.diffusion import create_diffusion
    sd, den = denoiser
    g = P.golden("sampler_L64_100")
    prot, z0, noises, steps = _sampler_inputs(g)
    diff = create_diffusion(str(steps))
    assert np.array_equal(np.array(diff.timestep_map), g["timestep_map"])
    plan = eng.Plan(den, 1, 1, prot.L, "fp32")
    plan.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([prot.L]), prot.restype_full[1:-1][None].int(),
                    torch.zeros(1, dtype=torch.int32))
    plan.set_schedule(diff.timestep_map, diff.coef_table())
    x = z0.cuda().clone()
    plan.sample(x, noises.cuda().contiguous(), use_graph=False)
    err = P.rel_err(x.cpu(), g["sample_0"])
    assert err < 1e-3, f"100-step latent vs reference: rel {err}"
    xg = z0.cuda().clone()
    plan.sample(xg, noises.cuda().contiguous(), use_graph=True)
    assert torch.equal(xg, x), "CUDA-graph replay must equal the eager launch sequence bit for bit"
    plan.sample(xg.copy_(z0), noises.cuda().contiguous(), use_graph=True)       # replay of the cached graph
    assert torch.equal(xg, x)


def test_generic_p_sample_matches_fused(eng, denoiser):
    """SpacedDiffusion.p_sample_loop_progressive with an arbitrary callable == fused plan.sample."""
    from codlad_b200.diffusion import create_diffusion
    sd, den = denoiser
    prot = synthetic.make_protein(32, 1, seed=5)
    diff = create_diffusion("10")
    plan = eng.Plan(den, 1, 2, 32, "fp32")
    plan.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([32]), prot.restype_full[1:-1][None].int(),
                    torch.zeros(2, dtype=torch.int32))
    plan.set_schedule(diff.timestep_map, diff.coef_table())
    z0 = synthetic.latent_noise((2, 32, 3), 1).cuda()
    noises = synthetic.latent_noise((10, 2, 32, 3), 2).cuda()
    fused = plan.sample(z0.clone(), noises, use_graph=False)
    plan2 = eng.Plan(den, 1, 2, 32, "fp32")
    plan2.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([32]), prot.restype_full[1:-1][None].int(),
                     torch.zeros(2, dtype=torch.int32))
    model = lambda x, t, **kw: plan2.forward(x, t.float())
    last = None
    for last in diff.p_sample_loop_progressive(model, z0.shape, z0.clone(), clip_denoised=False, device="cuda", step_noise=noises):
        pass
    assert (last["sample"] - fused).abs().max() < 1e-4 * fused.abs().max()


# --------------------------------------------------------------------------------------- VQ
def test_vq_indices_bit_exact(eng, R):
    sd = weights.init_vae_decode_state(0)
    mean, std = weights.LATENT_STATS[("N6", "PED")]
    vae = eng.VaeEngine(sd, mean, std)
    g = P.golden("vq_20000")
    n, seed = (int(v) for v in g["meta"])
    raw = synthetic.latent_noise((n, 3), seed)
    x = torch.tensor(mean) + torch.tensor(std) * raw
    ref = R.vq_nearest(x, sd["quantize._codebook.embed"][0])
    from codlad_b200 import _native as N
    NB, L = 100, 200
    lengths = torch.full((1,), L, dtype=torch.int32).cuda()
    frame_of = torch.zeros(NB, dtype=torch.int32).cuda()
    idx = torch.empty(NB, L, dtype=torch.int32, device="cuda")
    zq = torch.empty(NB, L, 3, device="cuda")
    xs = x.cuda().contiguous()
    N.check(N.lib().cb2_vq_lookup(vae.handle, N.dptr(xs), NB, L, N.dptr(lengths), N.dptr(frame_of), 0, N.dptr(idx), N.dptr(zq), N.stream_ptr()))
    assert torch.equal(idx.cpu().reshape(-1).long(), ref), "VQ code indices must be bit-exact vs the oracle"
    assert torch.equal(zq.cpu().reshape(-1, 3), sd["quantize._codebook.embed"][0][ref])
    assert (idx.cpu().reshape(-1).numpy() != g["idx"].reshape(-1)).mean() < 1e-3      # in-repo VectorQuantizerEMA (near-ties only)
    # fused de-normalisation (x*std + mean with two roundings) gives the same bits as the torch ops
    rs = raw.cuda().contiguous()
    idx2 = torch.empty_like(idx)
    N.check(N.lib().cb2_vq_lookup(vae.handle, N.dptr(rs), NB, L, N.dptr(lengths), N.dptr(frame_of), 1, N.dptr(idx2), N.dptr(zq), N.stream_ptr()))
    assert torch.equal(idx2, idx)
    # masked positions: index -1, input passed through
    short = torch.full((1,), 150, dtype=torch.int32).cuda()
    N.check(N.lib().cb2_vq_lookup(vae.handle, N.dptr(xs), NB, L, N.dptr(short), N.dptr(frame_of), 0, N.dptr(idx2), N.dptr(zq), N.stream_ptr()))
    assert bool((idx2[:, 150:] == -1).all()) and torch.equal(idx2[:, :150], idx[:, :150])
    assert torch.equal(zq[:, 150:].cpu(), x.reshape(NB, L, 3)[:, 150:])


# --------------------------------------------------------------------------------------- decode
def _decode_case(eng, g, angle, c2, ens=1):
    from codlad_b200 import sampler
    L, frames, prot_seed, lat_seed, _ = (int(v) for v in g["meta"])
    sd = P.decode_state(angle, c2)
    prot = synthetic.make_protein(L, frames, seed=prot_seed)
    batch = synthetic.collate(prot)
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    latent = mean + std * synthetic.latent_noise((frames, L, 3), lat_seed)
    fs = sampler.frames_from_batch(batch, prot.info, ens)
    return sd, prot, batch, latent, fs


@pytest.mark.parametrize("name,angle,c2", [("decode_L64_N6", False, False), ("decode_L64_K4", True, False),
                                           ("decode_L64_N6_c2", False, True), ("decode_c2_L300_N6", False, False),
                                           ("decode_c4_L2000_K4", True, False)])
def test_decode_vs_oracle_and_golden(eng, R, denoiser, name, angle, c2):
    _, den = denoiser
    g = P.golden(name)
    sd, prot, batch, latent, fs = _decode_case(eng, g, angle, c2)
    mean, std = weights.LATENT_STATS[("N6", "PED")]
    vae = eng.VaeEngine(sd, mean, std, angle)
    plan = eng.Plan(den, fs.F, fs.NB, fs.L, "fp32")
    plan.set_frames(fs.X, fs.lengths, fs.cg_z, fs.frame_of)
    plan.set_topology(vae, fs.ca_full, fs.csr_row, fs.csr_col, fs.orders, fs.slot_atom, fs.out_off)
    idx, zq, ic, xyz = plan.decode(vae, latent.cuda(), denorm=False, num_atoms_total=fs.total_atoms)
    mask = torch.ones(fs.F, fs.L, dtype=torch.bool)
    ic_ref, idx_ref = R.latent_decode(sd, latent, mask, batch["CG_nxyz"][:, 0].long(), batch["CG_nxyz"][:, 1:],
                                      batch["CG_nbr_list"], batch["num_CGs"], angle)
    assert torch.equal(idx.cpu().long(), idx_ref)
    ic = ic.cpu().reshape(-1, 13, 3)
    assert P.rel_err(ic, ic_ref) < 1e-5 and P.rel_err(ic, g["ic_recon"]) < 1e-5
    if not c2:
        xyz = xyz.cpu().reshape(fs.F, -1, 3)
        r = P.rmsd(xyz, g["xyz"])
        assert r < 1e-3, f"coordinates vs reference golden: RMSD {r} A"


def test_ic_to_xyz_large_angles(eng):
    from codlad_b200 import sampler, _native as N
    g = P.golden("ic_large_angle_L48")
    L, seed = (int(v) for v in g["meta"])
    prot = synthetic.make_protein(L, 2, seed=seed)
    fs = sampler.frames_from_batch(synthetic.collate(prot), prot.info, 1)
    ic = torch.from_numpy(g["ic"]).cuda().contiguous()
    xyz = torch.zeros(fs.total_atoms, 3, device="cuda")
    dev = [t.cuda() for t in (fs.ca_full, fs.frame_of, fs.lengths, fs.orders, fs.slot_atom, fs.out_off)]   # keep alive across the call
    N.check(N.lib().cb2_ic_to_xyz(N.dptr(dev[0]), N.dptr(ic), fs.NB, L, N.dptr(dev[1]), N.dptr(dev[2]),
                                  N.dptr(dev[3]), N.dptr(dev[4]), N.dptr(dev[5]), N.dptr(xyz), N.stream_ptr()))
    r = P.rmsd(xyz.cpu().reshape(2, -1, 3), g["xyz"])
    assert r < 1e-3, f"RMSD {r} A"


# --------------------------------------------------------------------------------------- whole path, config 1
def test_full_path_config1_fp32(eng, R):
    """Config 1 (1 x 64 residues, ensemble 1, 100 steps): stage-wise parity of the whole path."""
    from codlad_b200 import sampler
    dsd = weights.init_denoiser_state(0)
    vsd = weights.init_vae_decode_state(0)
    bm = sampler.Backmapper(dsd, vsd, "N6", precision="fp32")
    prot = synthetic.make_protein(64, 1, seed=1001)
    batch = synthetic.collate(prot)
    fs = sampler.frames_from_batch(batch, prot.info, 1)
    z0 = synthetic.latent_noise((1, 64, 3), 2001)
    noises = synthetic.latent_noise((100, 1, 64, 3), 3001)
    plan = bm.upload(fs)
    out = bm.sample(plan, fs, z0, noises)
    g = P.golden("sampler_L64_100")
    assert P.rel_err(out["latent"].cpu(), g["sample_0"]) < 1e-3
    # decode stages on the GPU latent (identical input to both sides)
    lat = out["latent"].cpu()
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    den = lat * std + mean
    mask = torch.ones(1, 64, dtype=torch.bool)
    ic_ref, idx_ref = R.latent_decode(vsd, den, mask, batch["CG_nxyz"][:, 0].long(), batch["CG_nxyz"][:, 1:],
                                      batch["CG_nbr_list"], batch["num_CGs"], False)
    assert torch.equal(out["idx"].cpu().long(), idx_ref)
    xyz_ref = R.ic_to_xyz(batch["OG_CG_nxyz"].reshape(-1, 66, 4), ic_ref.reshape(1, 64, 13, 3), prot.info)
    assert P.rmsd(out["xyz"].cpu().reshape(1, -1, 3), xyz_ref) < 1e-3


# --------------------------------------------------------------------------------------- reference call surface (drop-in)
def test_reference_call_surface_end_to_end(R):
    """The path exactly as test.py drives it (test.py:495-582): model factory + load_state_dict, create_diffusion,
    doubled batch cat([z, z]), p_sample_loop(model.forward, ...), chunk(2)[0], get_norm_feature(norm_in=False),
    latent_decode, ic_to_xyz -- each stage against the CPU oracle."""
    from codlad_b200.diffusion import create_diffusion
    from codlad_b200.latent_model import MPNN_models
    from codlad_b200.utils_ic import ic_to_xyz
    from codlad_b200.vae_model import VAE, get_norm_feature
    dsd, vsd = weights.init_denoiser_state(3), weights.init_vae_decode_state(3)
    model = MPNN_models["mpnn_diffusion"](input_size=3, unconditional=True, diffusion="diffusion", precision="fp32")
    assert list(model.state_dict().keys()) == list(dsd.keys())
    model.load_state_dict({"module." + k: v for k, v in dsd.items()})          # DDP-prefixed checkpoint, test.py:277-285
    vae = VAE("N6")
    vae.load_state_dict(vsd)
    T, L, Fr = 20, 56, 2
    diffusion = create_diffusion(str(T))
    prot = synthetic.make_protein(L, Fr, seed=4242)
    batch = synthetic.collate(prot)
    mask = torch.ones(Fr, L, dtype=torch.bool)
    z = synthetic.latent_noise((Fr, L, 3), 7)
    noises = synthetic.latent_noise((T, Fr, L, 3), 8)
    cat_z, cat_mask = torch.cat([z, z], 0), torch.cat([mask, mask], 0)
    samples = diffusion.p_sample_loop(model.forward, cat_z.shape, cat_z.cuda(), clip_denoised=False,
                                      model_kwargs=dict(y=None, mask=cat_mask.cuda(), batch=batch), device="cuda",
                                      step_noise=torch.cat([noises, noises], 1))
    first, second = samples.chunk(2, dim=0)
    assert torch.equal(first, second)                                          # both halves of the doubled batch are the same sample
    X = prot.ca_full[:, 1:-1].contiguous()
    zz = prot.restype_full[1:-1][None].expand(Fr, -1)
    ref_lat = R.sample_loop(dsd, z, X, zz, mask, noises, R.respaced_schedule(T))
    assert P.rel_err(first.cpu(), ref_lat) < 1e-4
    lat = get_norm_feature(first, "N6", norm_channel=True, norm_single=False, norm_in=False, dataname="PED")   # test.py:548
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    assert torch.equal(lat.cpu(), first.cpu() * std + mean)
    zq, idx, loss = vae.quantize(lat, mask=mask.cuda())
    _, ic_recon = vae.latent_decode(lat, mask.cuda(), batch)
    ic_ref, idx_ref = R.latent_decode(vsd, lat.cpu(), mask, batch["CG_nxyz"][:, 0].long(), batch["CG_nxyz"][:, 1:], batch["CG_nbr_list"],
                                      batch["num_CGs"], False)
    assert torch.equal(idx.cpu(), idx_ref) and float(loss) == 0.0
    assert ic_recon.shape == ic_ref.shape and P.rel_err(ic_recon.cpu(), ic_ref) < 1e-5
    og = batch["OG_CG_nxyz"].reshape(-1, L + 2, 4)
    xyz = ic_to_xyz(og, ic_recon.reshape(-1, L, 13, 3), prot.info)
    xyz_ref = R.ic_to_xyz(og, ic_ref.reshape(-1, L, 13, 3), prot.info)
    assert xyz.shape == xyz_ref.shape and P.rmsd(xyz.cpu(), xyz_ref) < 1e-3


def test_module_forward_matches_oracle_ragged_batch(R):
    """forward(x, t, y, mask, batch) on a ragged two-protein batch (padding + mask), fp32 tier."""
    from codlad_b200.latent_model import MPNN_models
    sd = weights.init_denoiser_state(0)
    model = MPNN_models["mpnn_diffusion"](precision="fp32")
    model.load_state_dict(sd)
    g = P.golden("denoiser_ragged")
    c = P.denoiser_case(g["meta"], [80, 70])
    cg = torch.cat([torch.cat([c["cg_z"][b, :n, None].float(), c["X"][b, :n]], 1) for b, n in enumerate([80, 70])], 0)
    batch = {"CG_nxyz": cg, "num_CGs": torch.tensor([80, 70])}
    out = model(c["x"].cuda(), c["t"].cuda(), None, mask=c["mask"].cuda(), batch=batch).cpu()
    m = c["mask"] & torch.tensor([True, True])[:, None]
    m[1] = False                                       # rows of the shorter protein see topk's arbitrary tie order in the reference
    assert np.abs(out.numpy() - g["out"])[m.numpy()].max() < 5e-5


# --------------------------------------------------------------------------------------- parity AT the BASELINE.json shapes
def _members_plan(eng, c, precision):
    den = eng.DenoiserEngine(weights.init_denoiser_state(0), c["k_neighbors"])
    plan = eng.Plan(den, 1, c["members"], c["L"], precision)
    plan._keep = den
    plan.set_frames(c["X"], torch.tensor([c["L"]]), c["cg_z"].int(), torch.zeros(c["members"], dtype=torch.int32))
    return plan


@pytest.mark.parametrize("name", ["denoiser_c2_L300x10", "denoiser_c4_L2000_K48x2"])
@pytest.mark.parametrize("precision,rel_bar,abs_bar", [("fp32", 5e-6, 2e-5), ("f16", 1.5e-3, 6e-3)])
def test_denoiser_forward_at_baseline_shapes_vs_reference(eng, name, precision, rel_bar, abs_bar):
    """configs[1] (one 300-residue frame x 10 members, t from 999 to 0) and configs[3] (2000 residues, k = 48, 2 members): one
    forward of BOTH tiers against the output of the unmodified reference (tests/golden, oracle/make_goldens.py).
    Bars (~3x measured): fp32 tier 5e-6 relative / 2e-5 max-abs (measured 1e-6 / 6e-6); f16 tier 1.5e-3 relative / 6e-3 max-abs on
    O(1) outputs (measured 4.8e-4 / 2.5e-3)."""
    g = P.golden(name)
    c = P.members_case(g["meta"])
    plan = _members_plan(eng, c, precision)
    K = plan.K
    assert P.knn_tie_aware_equal(plan.buffer("nbr_idx").cpu().numpy().reshape(-1, K)[::8], g["E_idx"].reshape(-1, K).astype(np.int64)[::8],
                                 g["D_nb"].reshape(-1, K))
    out = plan.forward(c["x"].cuda(), c["t"].float().cuda()).cpu()
    ref = torch.from_numpy(g["out"])
    rel, mx = P.rel_err(out, ref), float((out - ref).abs().max())
    print(f"{name} {precision}: rel {rel:.3e} max-abs {mx:.3e}")
    assert rel < rel_bar and mx < abs_bar


@pytest.mark.parametrize("precision,rel_bar,abs_bar", [("fp32", 5e-6, 2e-5), ("f16", 1.5e-3, 6e-3)])
def test_denoiser_forward_c3_shape_vs_reference(eng, denoiser, precision, rel_bar, abs_bar):
    """configs[2]-shaped: three 500-residue frames, one member each."""
    _, den = denoiser
    g = P.golden("denoiser_c3_L500_B3")
    c = P.denoiser_case(g["meta"])
    plan = _plan_for_case(eng, den, c, precision=precision, keep_debug=False)
    out = plan.forward(c["x"].cuda(), c["t"].float().cuda()).cpu()
    ref = torch.from_numpy(g["out"])
    rel, mx = P.rel_err(out, ref), float((out - ref).abs().max())
    print(f"denoiser_c3_L500_B3 {precision}: rel {rel:.3e} max-abs {mx:.3e}")
    assert rel < rel_bar and mx < abs_bar


@pytest.mark.parametrize("precision,bar", [("fp32", 5e-6), ("f16", 1.5e-3)])       # measured 3.7e-7 / 4.3e-4
def test_sampler_5_steps_at_c2_shape_vs_reference(eng, precision, bar):
    """5-step p_sample_loop of 10 members on the 300-residue frame (graph and eager) against the reference's own loop."""
    from codlad_b200.diffusion import create_diffusion
    g = P.golden("sampler_c2_L300x10_5")
    L, members, prot_seed, z_seed, noise_seed, steps = (int(v) for v in g["meta"])
    c = dict(P.members_case([L, members, 64, prot_seed, 0]))
    plan = _members_plan(eng, c, precision)
    diff = create_diffusion(str(steps))
    assert np.array_equal(np.array(diff.timestep_map), g["timestep_map"])
    plan.set_schedule(diff.timestep_map, diff.coef_table())
    z0 = synthetic.latent_noise((members, L, 3), z_seed).cuda()
    nz = synthetic.latent_noise((steps, members, L, 3), noise_seed).cuda().contiguous()
    x = plan.sample(z0.clone(), nz, use_graph=False)
    err = P.rel_err(x.cpu(), g["sample_0"])
    print(f"sampler_c2_L300x10_5 {precision}: rel {err:.3e}")
    assert err < bar
    xg = plan.sample(z0.clone(), nz, use_graph=True)
    assert torch.equal(xg, x)


# --------------------------------------------------------------------------------------- plan re-use hazards
def test_module_plan_cache_follows_batch_content():
    """Two different batches of the same shape in sequence, the first freed (the reference driver loop, test.py:481-534): the
    module must not serve the second batch from the plan built for the first (round-1 bug: cache keyed on data_ptr)."""
    import gc
    from codlad_b200.latent_model import MPNN_models
    model = MPNN_models["mpnn_diffusion"](precision="fp32")
    model.load_state_dict(weights.init_denoiser_state(0))
    L = 40
    x = synthetic.latent_noise((1, L, 3), 3).cuda()
    t = torch.tensor([500.0]).cuda()
    mask = torch.ones(1, L, dtype=torch.bool).cuda()
    outs, solo = [], []
    for seed in (31, 32, 33, 34):
        batch = {k: v.cuda() if torch.is_tensor(v) else v for k, v in synthetic.collate(synthetic.make_protein(L, 1, seed=seed)).items()}
        outs.append(model(x, t, None, mask=mask, batch=batch).cpu())
        del batch
        gc.collect()
        torch.cuda.empty_cache()
    for seed in (31, 32, 33, 34):
        fresh = MPNN_models["mpnn_diffusion"](precision="fp32")
        fresh.load_state_dict(weights.init_denoiser_state(0))
        solo.append(fresh(x, t, None, mask=mask, batch=synthetic.collate(synthetic.make_protein(L, 1, seed=seed))).cpu())
    for a, b in zip(outs, solo):
        assert torch.equal(a, b)
    assert not torch.equal(outs[0], outs[1])
    # in-place edit of the coordinates of a live batch is seen too (version counter / content check)
    batch = synthetic.collate(synthetic.make_protein(L, 1, seed=31))
    a = model(x, t, None, mask=mask, batch=batch).cpu()
    batch["CG_nxyz"][:, 1:] += 0.25 * torch.randn(L, 3, generator=torch.Generator().manual_seed(1))
    assert not torch.equal(model(x, t, None, mask=mask, batch=batch).cpu(), a)


def test_graph_follows_the_mask_geometry(eng, denoiser):
    """A full-length frame set, then a ragged one of the same padded shape, then full again, sampled through ONE plan with
    use_graph=True: the graph must be re-captured when the masked / unmasked kernel choice changes (round-1 bug: the
    unmasked message kernel was replayed on ragged frames)."""
    from codlad_b200.diffusion import create_diffusion
    _, den = denoiser
    L, T = 72, 6
    diff = create_diffusion(str(T))
    full = P.denoiser_case([L, 2, 64, 1501, 1, 0, 0])
    ragged = P.denoiser_case([L, 0, 64, 1502, 1, 0, 0], [L, 50])
    z0 = synthetic.latent_noise((2, L, 3), 4).cuda()
    nz = synthetic.latent_noise((T, 2, L, 3), 5).cuda().contiguous()
    plan = eng.Plan(den, 2, 2, L, "f16")
    plan.set_schedule(diff.timestep_map, diff.coef_table())
    x, noise = torch.empty_like(z0), nz
    res = {}
    for tag, c in (("full", full), ("ragged", ragged), ("full2", full)):
        plan.set_frames(c["X"], c["mask"].sum(1).int(), c["cg_z"].int(), torch.arange(2, dtype=torch.int32))
        x.copy_(z0)
        plan.sample(x, noise, use_graph=True)
        res[tag] = x.clone()
        fresh = eng.Plan(den, 2, 2, L, "f16")
        fresh.set_schedule(diff.timestep_map, diff.coef_table())
        fresh.set_frames(c["X"], c["mask"].sum(1).int(), c["cg_z"].int(), torch.arange(2, dtype=torch.int32))
        want = fresh.sample(z0.clone(), nz, use_graph=False)
        valid = c["mask"].cuda()
        assert torch.equal(res[tag][valid], want[valid]), tag
    assert torch.equal(res["full"], res["full2"])


# --------------------------------------------------------------------------------------- BASELINE configs[2] / [3] shapes
@pytest.mark.parametrize("L,K,frames,ens", [(2000, 48, 1, 4), (500, 64, 6, 1)])
def test_whole_path_at_large_config_shapes(eng, L, K, frames, ens):
    """configs[3]-shaped (one 2000-residue frame, k_neighbors=48, K4 decoder) and configs[2]-shaped (several 500-residue
    proteins, K3 decoder) passes through the f16 tier: many tiles per CTA and a working set beyond L2.  The oracle
    would need hours here, so the checks are the size-independent ones: the CUDA-graph replay equals the eager loop bit
    for bit (the step loop has a fixed reduction order), outputs are finite, ensemble members that share a frame and
    start from the same noise give identical coordinates, and the decoded bond lengths are the residue-type table
    entries (a property of the IC decoder whatever the latent).  Regression test for the barrier-parity hazard that
    dead-locked the per-edge kernel on exactly this kind of working set."""
    from codlad_b200 import sampler
    vae_type = "K4" if frames == 1 else "K3"
    prots = [synthetic.make_protein(L, 1, seed=4100 + i) for i in range(frames)]
    batch = synthetic.collate(prots[0]) if frames == 1 else synthetic.collate_many(prots)
    fs = sampler.frames_from_batch(batch, [p.info for p in prots], ens)
    vsd = weights.init_vae_decode_state(0, True, (vae_type, sampler.VAE_DATA[vae_type]))
    bm = sampler.Backmapper(weights.init_denoiser_state(0), vsd, vae_type, k_neighbors=K, precision="f16")
    plan = bm.upload(fs)
    z = synthetic.latent_noise((fs.F, L, 3), 7).repeat(ens, 1, 1)                       # members of a frame share their noise
    noise = synthetic.latent_noise((bm.T, fs.F, L, 3), 8).repeat(1, ens, 1, 1)
    eager = {k: v.clone() for k, v in bm.sample(plan, fs, z, noise, use_graph=False).items() if v is not None}
    graph = bm.sample(plan, fs, z, noise, use_graph=True)
    for k in ("latent", "xyz", "ic_recon"):
        assert torch.isfinite(eager[k]).all(), k
        assert torch.equal(eager[k], graph[k]), f"{k}: graph replay differs from the eager loop"
    lat = eager["latent"].reshape(ens, fs.F, L, 3)
    assert torch.equal(lat[0], lat[-1]), "members with identical noise diverged"
    ic = eager["ic_recon"].reshape(fs.NB, L, 13, 3).cpu()
    table = vsd["equivaraintconv.backbone_dist.weight"]
    cg_z = fs.cg_z.long()[fs.frame_of.long()]
    assert torch.equal(ic[:, :, :3, 0], table[cg_z]), "backbone bond lengths are a table lookup"


def test_training_losses_with_cuda_denoiser(R):
    """Validation loss as train_latent computes it (train_latent.py:208: diffusion.training_losses(model, x, t, model_kwargs)),
    with the CUDA denoiser behind the reference module surface, against the same formulas fed with the oracle's forward."""
    from codlad_b200.diffusion import create_diffusion
    from codlad_b200.latent_model import MPNN_models
    dsd = weights.init_denoiser_state(5)
    model = MPNN_models["mpnn_diffusion"](input_size=3, unconditional=True, diffusion="diffusion", precision="fp32")
    model.load_state_dict(dsd)
    diffusion = create_diffusion("")
    L, Fr = 48, 3
    prot = synthetic.make_protein(L, Fr, seed=4343)
    batch = synthetic.collate(prot)
    mask = torch.ones(Fr, L, dtype=torch.bool)
    x0 = synthetic.latent_noise((Fr, L, 3), 17)
    noise = synthetic.latent_noise((Fr, L, 3), 18)
    t = torch.tensor([0, 412, 999])
    terms = diffusion.training_losses(model.forward, x0.cuda(), t.cuda(), dict(y=None, mask=mask.cuda(), batch=batch), noise=noise.cuda())
    X = prot.ca_full[:, 1:-1].contiguous()
    zz = prot.restype_full[1:-1][None].expand(Fr, -1)
    ref_model = lambda x_t, tt, **kw: R.denoiser_forward(dsd, x_t, tt, X, zz, mask, 64)
    ref = diffusion.training_losses(ref_model, x0, t, dict(mask=mask), noise=noise)
    for k in ("mse", "vb", "loss"):
        assert torch.allclose(terms[k].cpu(), ref[k], rtol=1e-3, atol=1e-5), (k, terms[k], ref[k])


# --------------------------------------------------------------------------------------- evaluation step
@pytest.mark.parametrize("name", ["sample_qualities_exact", "sample_qualities_noisy"])
def test_bond_graph_metrics_bit_exact_and_vs_reference_golden(R, name):
    """cb2_eval_bond_graphs: the six integer counts per structure equal the oracle's exactly, the derived metrics equal what the
    unmodified reference's eval_sample_qualities returned (golden)."""
    from codlad_b200 import metrics
    g = P.golden(name)
    ref, gen = torch.from_numpy(g["xyz_ref"]), torch.from_numpy(g["xyz_gen"])
    z, num = torch.from_numpy(g["z"].astype(np.int64)), g["num_atoms"].tolist()
    counts, sums = metrics.bond_graph_stats(ref.cuda(), gen.cuda(), z.cuda(), num)
    c_ref, s_ref = R.sample_quality_stats(ref, gen, z, num)
    assert torch.equal(counts.cpu(), c_ref)
    assert torch.allclose(sums.cpu(), s_ref, rtol=1e-12, atol=0)
    q = metrics.eval_sample_qualities(ref.cuda(), gen.cuda(), z.cuda(), num)
    want = g["q"]
    assert (q["heavy_valid"].cpu().double().numpy() == want[:, 0]).all() and (q["all_valid"].cpu().double().numpy() == want[:, 1]).all()
    np.testing.assert_allclose(q["heavy_graph_diff_ratio"].cpu().numpy(), want[:, 2], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(q["all_graph_diff_ratio"].cpu().numpy(), want[:, 3], rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(q["all_rmsd"].cpu().numpy(), want[:, 4], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(q["heavy_rmsd"].cpu().numpy(), want[:, 5], rtol=1e-6, atol=1e-9)
    lists = metrics.valid_ratio_and_cut_off_result(ref.cuda(), gen.cuda(), torch.tensor(num), z)
    assert lists[0] == want[:, 0].tolist() and lists[1] == want[:, 1].tolist() and len(lists[2]) == len(num)


def test_bond_graph_metrics_on_backmapped_ensemble(eng, R):
    """The metric on what the path produces: a 120-residue protein backmapped 4 times (f16 tier), every member scored against
    member 0 -- ragged structure sizes are exercised by mixing in a second protein; counts equal the oracle's exactly."""
    from codlad_b200 import metrics, sampler
    prots = [synthetic.make_protein(120, 1, seed=911), synthetic.make_protein(77, 1, seed=912)]
    batch = synthetic.collate_many(prots)
    fs = sampler.frames_from_batch(batch, [p.info for p in prots], 2)
    bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0), "N6", num_sampling_steps=10, precision="f16")
    out = bm.sample(bm.upload(fs), fs, generator=torch.Generator(device="cuda").manual_seed(3))
    xyz = out["xyz"]
    num = fs.num_atoms.repeat(2).tolist()                       # member order: ensemble-major (b = e * F + f)
    first = xyz[:sum(num[:2])]
    ref = torch.cat([first, first], 0)
    z = torch.tensor([7, 6, 6, 8, 6, 6, 16, 8, 1], dtype=torch.int64).repeat(xyz.shape[0] // 9 + 1)[:xyz.shape[0]]
    counts, sums = metrics.bond_graph_stats(ref, xyz, z.cuda(), num)
    c_ref, s_ref = R.sample_quality_stats(ref.cpu(), xyz.cpu(), z, num)
    assert torch.equal(counts.cpu(), c_ref)
    assert torch.allclose(sums.cpu(), s_ref, rtol=1e-12, atol=0)
    assert int(counts[0, 0]) == 0 and int(counts[1, 0]) == 0    # members 0 are scored against themselves


def test_superposed_rmsd_and_diversity_vs_oracle(R):
    """cb2_superposed_rmsd / metrics.compute_div (test.py:37-96) against the float64 Kabsch restatement: rotated + translated copies give
    0, noisy copies agree to 1e-6 relative, reflections are not allowed (a mirrored structure keeps a finite RMSD), ragged sizes."""
    from codlad_b200 import metrics
    g = torch.Generator().manual_seed(9)
    sizes = [517, 60, 2479, 3]
    A = [torch.randn(n, 3, generator=g) * 8 for n in sizes]
    q = torch.randn(4, generator=g); q = q / q.norm()
    w, x, y, z = q.tolist()
    Rm = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    B = [a @ Rm.T + torch.tensor([3.0, -7.0, 11.0]) + (0.0 if k == 0 else 0.3) * torch.randn(a.shape, generator=g) for k, a in enumerate(A)]
    B[1] = A[1] * torch.tensor([1.0, 1.0, -1.0])                      # mirror image
    got = metrics.superposed_rmsd(torch.cat(A), torch.cat(B), sizes).cpu()
    want = torch.tensor([R.superposed_rmsd(a.numpy(), b.numpy()) for a, b in zip(A, B)], dtype=torch.float64)
    assert float(got[0]) < 1e-5 and float(want[1]) > 1.0
    assert torch.allclose(got[1:], want[1:], rtol=1e-6, atol=1e-9), (got, want)
    gen = [torch.randn(3, 200, 3, generator=g) * 5 for _ in range(4)]
    ref = torch.randn(3, 200, 3, generator=g) * 5
    div, r_ref, r_gen = metrics.compute_div(gen, ref)
    d0, a0, b0 = R.compute_div([v.numpy() for v in gen], ref.numpy())
    assert abs(div - d0) < 1e-6 and abs(r_ref - a0) < 1e-5 * a0 and abs(r_gen - b0) < 1e-5 * b0


def test_pair_list_losses_vs_reference(R):
    """cb2_pair_losses / cb2_keys_once behind metrics.{inter,clash,ged}_result against the unmodified reference functions
    (test.py:97-146; tests/golden/eval_losses.npz) and the oracle restatement.  The once-only pair selection is exact (duplicates
    inside one list, pairs present in both lists, reversed pairs are distinct rows); the violation count is an integer and agrees
    exactly; means agree to the reference's fp32 summation error (the kernel sums in double)."""
    from codlad_b200 import metrics
    gold = P.golden("eval_losses")
    c = synthetic.eval_loss_case(int(gold["meta"][0]))
    want = gold["vals"]
    na = c["xyz"].shape[0]
    cu = {k: v.cuda() for k, v in c.items()}
    want_pairs = R.clash_pairs(c["edge"], c["nbr"])
    got_pairs = metrics.rows_once(cu["edge"], cu["nbr"], na).cpu()
    assert torch.equal(got_pairs, want_pairs) and got_pairs.shape[0] == int(want[5])
    o = metrics.pair_losses(cu["recon"], want_pairs.cuda(), thr_count=1.2).cpu()
    d = ((c["recon"][want_pairs[:, 0]] - c["recon"][want_pairs[:, 1]]).pow(2).sum(-1) + 1e-7).sqrt()
    assert int(o[0]) == int((d < 1.2).sum()) and int(o[0]) > 10 and int(o[3]) == want_pairs.shape[0]
    close = lambda a, b: abs(float(a) - float(b)) <= 3e-6 * max(abs(float(b)), 1e-3)
    assert close(metrics.clash_result(cu["edge"], cu["nbr"], cu["recon"], cu["bb"]), want[3])
    assert close(metrics.ged_result(cu["recon"], cu["xyz"], cu["edge"]), want[4])
    gi, gp = metrics.inter_result(cu["inter"], cu["pipi"], cu["recon"])
    assert close(gi, want[0]) and close(gp, want[1])
    gi0, gp0 = metrics.inter_result(cu["inter"], cu["pipi"][:0], cu["recon"])
    assert close(gi0, want[2]) and float(gp0) == 0.0
    for il, pl in ((c["inter"][:0], c["pipi"]), (c["inter"][:0], c["pipi"][:0])):
        gi, gp = metrics.inter_result(il.cuda(), pl.cuda(), cu["recon"])
        wi, wp = R.inter_result(il, pl, c["recon"])
        assert close(gi, wi) and close(gp, wp)
    with pytest.raises(IndexError):
        metrics.pair_losses(cu["xyz"], torch.tensor([[0, na]]).cuda())


def test_repeated_frames_in_a_batch_share_their_precompute():
    """An ensemble written as repeated frames in the reference batch schema (and test.py's doubled batch) is served by a plan that holds the
    DISTINCT frames only; results equal those of one-frame batches row by row."""
    from codlad_b200.latent_model import MPNN_models
    model = MPNN_models["mpnn_diffusion"](precision="fp32")
    model.load_state_dict(weights.init_denoiser_state(0))
    L = 44
    p2 = synthetic.make_protein(L, 2, seed=808)                      # two conformations of one protein
    batch = synthetic.collate(p2, frames=[0, 1, 0, 0, 1])
    x = synthetic.latent_noise((5, L, 3), 4).cuda()
    t = torch.tensor([10.0, 500.0, 999.0, 10.0, 250.0]).cuda()
    mask = torch.ones(5, L, dtype=torch.bool).cuda()
    out = model(x, t, None, mask=mask, batch=batch).cpu()
    plan = model.plan_for(batch, 5)
    assert plan.F == 2 and plan.NB == 5
    for b, f in enumerate([0, 1, 0, 0, 1]):
        solo = MPNN_models["mpnn_diffusion"](precision="fp32")
        solo.load_state_dict(weights.init_denoiser_state(0))
        one = solo(x[b:b + 1], t[b:b + 1], None, mask=mask[:1], batch=synthetic.collate(p2, frames=[f])).cpu()
        assert torch.equal(out[b:b + 1], one), b


def test_sharded_backmapper_maps_every_unit_to_its_frame():
    """sampler.ShardedBackmapper on one rank (the multi-rank exchange is covered on CPU with gloo): ragged proteins are grouped into padded
    plans, every (frame, member) unit comes back with its frame's atom count, and the C-alpha atoms of each structure ARE the input trace
    of that frame (slot 3 of ic_to_xyz copies it, utils_ic.py:247-266) -- so a unit can not have been decoded on another frame's geometry."""
    from codlad_b200 import sampler
    lengths = [90, 41, 88, 64, 43]
    prots = [synthetic.make_protein(n, 1, seed=1200 + i) for i, n in enumerate(lengths)]
    batches, infos = [synthetic.collate(p) for p in prots], [p.info for p in prots]
    bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0), "N6", num_sampling_steps=8, precision="f16")
    sb = sampler.ShardedBackmapper(bm, 0, 1, max_frames=3, max_pad=0.12)
    groups = sb.local_groups(sb.plan(lengths, 2)[0], lengths)
    assert len(groups) >= 2 and sorted(f for g in groups for f in g[0]) == list(range(5))
    out = sb.backmap(batches, infos, 2, generator=torch.Generator(device="cuda").manual_seed(5))
    assert sorted(out) == [(f, e) for f in range(5) for e in range(2)]
    for (f, e), xyz in out.items():
        permute, atom_idx, _ = infos[f]
        assert xyz.shape == (permute.numel(), 3) and torch.isfinite(xyz).all()
        ca_rows = torch.nonzero(atom_idx[permute] % 14 == 3)[:, 0]
        assert ca_rows.numel() == lengths[f]
        assert torch.equal(xyz[ca_rows], prots[f].ca_full[0, 1:-1]), (f, e)
    assert not torch.equal(out[(0, 0)], out[(0, 1)])                      # members differ (their own noise)
    again = sb.backmap(batches, infos, 2, generator=torch.Generator(device="cuda").manual_seed(5))
    assert all(torch.equal(out[k], again[k]) for k in out)                # same seeds, same result


def test_random_geometry_sweep():
    """tools/fuzz_parity.py on 30 random geometries (1-4 frames of ragged length 2-700, 1-12 members, k in {16, 30, 48, 64}): the f16 tier
    tracks the fp32 tier within 4e-3 (worst measured over 260 cases: 8.7e-4), outputs finite, CUDA-graph replay of a 3-step loop
    bit-identical to the eager loop."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "30", "5"], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "30 cases, worst" in r.stdout
