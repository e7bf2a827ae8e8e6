"""CPU: host-side logic of the product package (schedule tables, topology tables, batch -> FrameSet
conversion, weight inventories) against the oracle / the committed goldens."""
import numpy as np
import torch

from codlad_b200 import diffusion, sampler, synthetic, topology, weights
from oracle import restate as R
from tests import parity_utils as P


def test_schedule_tables_match_oracle_and_reference_golden():
    d = diffusion.create_diffusion("100")
    s = R.respaced_schedule(100)
    assert np.array_equal(np.array(d.timestep_map), s["timestep_map"])
    assert np.array_equal(np.array(d.timestep_map), P.golden("sampler_L64_100")["timestep_map"])
    c = d.coef_table()
    for col, key in enumerate(["post_logvar_clipped", "log_betas", "sqrt_recip_ac", "sqrt_recipm1_ac", "post_coef1", "post_coef2"]):
        assert np.array_equal(c[:, col], s[key].astype(np.float32)), key
    assert c[0, 6] == 0 and (c[1:, 6] == 1).all()
    for n in ("1", "10", "250", "ddim50", "10,20"):
        dd = diffusion.create_diffusion(n)
        assert dd.num_timesteps == len(dd.timestep_map)
    assert sorted(diffusion.space_timesteps(1000, "100")) == R.kept_timesteps(100)


def test_weight_inventories():
    d = weights.denoiser_shapes()
    assert len(d) == 108 and sum(int(np.prod(s)) for s in d.values()) == 2449974
    assert sum(int(np.prod(s)) for s in weights.ic_decoder_shapes(False).values()) == 43394
    assert sum(int(np.prod(s)) for s in weights.ic_decoder_shapes(True).values()) == 51044
    c2 = P.golden("ic_decoder_c2")
    assert sorted(c2.files) == sorted(weights.ic_decoder_shapes(False))


def test_topology_tables():
    assert topology.NUM_RESTYPES == 22 and int(topology.ATOM_COUNT.max()) == 14
    for rid in range(22):
        n_side = int(topology.ATOM_COUNT[rid]) - 4
        for s in range(10):
            tri = topology.ORDERS[rid, s]
            if s < n_side:
                assert int(tri.max()) < 4 + s, "an atom may only be built from atoms placed before it"
            else:
                assert tri.tolist() == [0, 1, 2]
    rt = torch.tensor([3, 13, 2, 16])
    permute, atom_idx, orders = topology.build_info(rt)
    assert permute.numel() == 4 + 14 + 5 + 7 and orders.shape == (10, 4, 3)
    inv = topology.slot_to_atom_map((permute, atom_idx, orders), 4)
    assert (inv >= 0).sum() == permute.numel() and inv[4] == -1 and inv[14] == 4


def test_frames_from_batch_matches_oracle_graph():
    prot = synthetic.make_protein(50, 3, seed=9)
    batch = synthetic.collate(prot)
    fs = sampler.frames_from_batch(batch, prot.info, num_ensemble=2)
    assert fs.F == 3 and fs.NB == 6 and fs.L == 50 and fs.total_atoms == 6 * prot.num_atoms
    assert torch.equal(fs.X, prot.ca_full[:, 1:-1]) and torch.equal(fs.ca_full, prot.ca_full)
    assert fs.frame_of.tolist() == [0, 1, 2, 0, 1, 2]
    nbr = R.directed_edges(batch["CG_nbr_list"])
    want = set(map(tuple, nbr.tolist()))
    got = set()
    for r in range(fs.F * fs.L):
        f = r // fs.L
        for e in range(int(fs.csr_row[r]), int(fs.csr_row[r + 1])):
            got.add((r, f * fs.L + int(fs.csr_col[e])))
    assert got == want
    for r in range(fs.F * fs.L):
        seg = fs.csr_col[int(fs.csr_row[r]):int(fs.csr_row[r + 1])]
        assert bool((seg[1:] > seg[:-1]).all())


def test_frames_from_batch_ragged():
    prots = [synthetic.make_protein(n, 1, seed=30 + i) for i, n in enumerate((40, 25))]
    parts = [synthetic.collate(p) for p in prots]
    batch = {k: torch.cat([p[k] for p in parts]) for k in ("CG_nxyz", "OG_CG_nxyz", "num_CGs")}
    batch["CG_nbr_list"] = torch.cat([parts[0]["CG_nbr_list"], parts[1]["CG_nbr_list"] + 40])
    fs = sampler.frames_from_batch(batch, [p.info for p in prots], 1)
    assert fs.L == 40 and fs.lengths.tolist() == [40, 25]
    assert torch.equal(fs.X[1, :25], prots[1].ca_full[0, 1:-1]) and float(fs.X[1, 25:].abs().sum()) == 0
    assert int(fs.csr_row[40 + 25]) == int(fs.csr_row[-1])          # padded rows have no edges
    assert fs.out_off.tolist() == [0, prots[0].num_atoms]


def test_reference_surface_modules_have_reference_state_dict_keys():
    """Module construction is CPU-only; only forward needs the GPU.  Keys / shapes are the drop-in contract."""
    from codlad_b200 import weights
    from codlad_b200.latent_model import MPNN_models, fused_sampler_for
    from codlad_b200.vae_model import VAE, get_norm_feature
    m = MPNN_models["mpnn_diffusion"](input_size=3, unconditional=True, diffusion="diffusion", self_condition=False)
    want = weights.denoiser_shapes()
    sd = m.state_dict()
    assert list(sd.keys()) == list(want.keys()) and len(sd) == 108
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in want.items())
    assert sum(v.numel() for v in sd.values()) == 2449974                   # parameter count of the reference
    assert fused_sampler_for(m.forward) is not None and fused_sampler_for(lambda *a: None) is None
    for vt, angle in (("N6", False), ("K4", True)):
        v = VAE(vt)
        assert list(v.state_dict().keys()) == list(weights.vae_decode_shapes(angle).keys())
    import torch
    x = torch.randn(2, 5, 3)
    # reference call form (test.py:548 / train_latent.py:194): feature_type = VAE type, dataname = data set; IDRome_test_7 is remapped
    y = get_norm_feature(x, "K4", norm_channel=True, norm_single=False, norm_in=False, dataname="Atlas")
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("K4", "Atlas")])
    assert torch.equal(y, x * std + mean)
    assert torch.allclose(get_norm_feature(y, "K4", norm_in=True, dataname="Atlas"), x, atol=1e-5)
    assert torch.equal(get_norm_feature(x, "N6", norm_in=False, dataname="IDRome_test_7"), get_norm_feature(x, "N6", norm_in=False, dataname="PED"))
    import pytest
    with pytest.raises(FileNotFoundError):
        get_norm_feature(x, "PED", norm_in=False, dataname="PED")          # the reference would fail to open PED_PED_x_mean.pt
    with pytest.raises(NotImplementedError):
        get_norm_feature(x, "N6", norm_single=True, norm_in=False, dataname="PED")
    # VAE.load_state_dict is strict on the decode-side tensors (a mismatched checkpoint must not decode with random weights),
    # ignores encoder extras and strips a DDP prefix
    v = VAE("N6")
    good = {"module." + k: t.clone() for k, t in weights.init_vae_decode_state(1).items()}
    good["module.encoder.some.weight"] = torch.zeros(3)
    v.load_state_dict(good)
    assert torch.equal(v.state_dict()["map_out.weight"], good["module.map_out.weight"])
    bad = dict(good)
    bad.pop("module.map_out.weight")
    with pytest.raises(RuntimeError):
        v.load_state_dict(bad)
    v.load_state_dict(bad, strict=False)


def test_batching_matches_reference_padding_and_make_directed():
    """pad_frames == reshape_and_create_mask's padding (models/gcn_nn.py:35-43); batch_csr == make_directed (:54-64) grouped by
    padded source row with frame-local, ascending neighbour ids."""
    import torch
    from codlad_b200 import batching, synthetic
    prots = [synthetic.make_protein(n, 1, seed=70 + i) for i, n in enumerate([12, 30, 7])]
    batch = synthetic.collate_many(prots)
    num = batch["num_CGs"]
    L = int(num.max())
    X, z = batching.pad_frames(batch["CG_nxyz"], num)
    ref = torch.nn.utils.rnn.pad_sequence(torch.split(batch["CG_nxyz"], num.tolist(), dim=0), batch_first=True)
    assert torch.equal(X, ref[..., 1:]) and torch.equal(z.float(), ref[..., 0])
    row_ptr, col = batching.batch_csr(batch["CG_nbr_list"], num, L)
    nbr = torch.cat([batch["CG_nbr_list"], batch["CG_nbr_list"].flip(1)], 0)          # undirected (i<j) input -> both directions
    off = torch.cumsum(num, 0) - num
    want = {}
    for a, b in nbr.tolist():
        f = int(torch.bucketize(torch.tensor(a), torch.cumsum(num, 0), right=True))
        want.setdefault(f * L + a - int(off[f]), []).append(b - int(off[f]))
    for r in range(len(prots) * L):
        got = col[int(row_ptr[r]):int(row_ptr[r + 1])].tolist()
        assert got == sorted(want.get(r, [])), r
    # an already directed list is taken as is
    rp2, col2 = batching.batch_csr(nbr, num, L)
    assert torch.equal(rp2, row_ptr) and torch.equal(col2, col)
    assert batching.batch_csr(torch.zeros(0, 2, dtype=torch.int64), num, L)[0].numel() == len(prots) * L + 1


def test_p_mean_variance_matches_oracle_update():
    """p_mean_variance (gaussian_diffusion.py:262-360) against the oracle's one-step update: mean + exp(log_variance / 2) * noise
    must reproduce oracle.restate.p_sample_update for every respaced step (pure host/torch glue, runs without the CUDA library)."""
    from codlad_b200.diffusion import create_diffusion
    from oracle import restate as R
    diff = create_diffusion("100")
    sched = R.respaced_schedule(100)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 7, 3, generator=g)
    out = torch.randn(2, 7, 6, generator=g)
    noise = torch.randn(2, 7, 3, generator=g)
    model = lambda x_, t_, **kw: out
    for step in (0, 1, 50, 99):
        t = torch.full((2,), step, dtype=torch.long)
        pmv = diff.p_mean_variance(model, x, t, clip_denoised=False)
        nz = 0.0 if step == 0 else 1.0
        got = pmv["mean"] + nz * torch.exp(0.5 * pmv["log_variance"]) * noise
        ref = R.p_sample_update(x, out, step, noise, sched)
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6), step
        assert torch.allclose(pmv["variance"], torch.exp(pmv["log_variance"]))


def test_training_losses_values_match_reference_golden():
    """training_losses / q_sample / posterior / VB terms against the golden written by the unmodified reference
    (oracle/make_goldens.py::golden_training_losses): loss values only, the package has no backward pass."""
    import os
    from codlad_b200.diffusion import create_diffusion
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "training_losses_B6.npz"))
    B, L, seed = (int(v) for v in g["meta"])
    diff = create_diffusion(timestep_respacing="")
    assert diff.num_timesteps == 1000
    gen = torch.Generator().manual_seed(seed)
    x_start = torch.randn(B, L, 3, generator=gen)
    noise = torch.randn(B, L, 3, generator=gen)
    model_out = 0.7 * torch.randn(B, L, 6, generator=gen)
    t = torch.tensor([0, 1, 500, 999, 37, 0][:B])
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[1, L - 5:] = False
    mask[-1, L // 2:] = False
    model = lambda x_t, tt, **kw: model_out
    terms = diff.training_losses(model, x_start, t, dict(mask=mask), noise=noise)
    for k in ("mse", "vb", "loss"):
        assert torch.allclose(terms[k], torch.from_numpy(g[k]), rtol=2e-5, atol=1e-6), k
    nomask = diff.training_losses(model, x_start, t, {}, noise=noise)
    assert torch.allclose(nomask["loss"], torch.from_numpy(g["loss_nomask"]), rtol=2e-5, atol=1e-6)


def test_oracle_superposed_rmsd_known_answers():
    """The float64 Kabsch restatement of md.rmsd (oracle, test infrastructure): invariance under rotation / translation, no reflection,
    and a closed-form case (two points at distance 2 vs distance 4 -> rmsd 1 after centring)."""
    import numpy as np
    from oracle import restate as R
    rng = np.random.default_rng(0)
    a = rng.normal(size=(50, 3))
    th = 0.7
    rot = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    assert R.superposed_rmsd(a, a @ rot.T + 5.0) < 1e-12
    assert R.superposed_rmsd(a, a * np.array([1, 1, -1])) > 0.1
    assert abs(R.superposed_rmsd(np.array([[-1.0, 0, 0], [1.0, 0, 0]]), np.array([[0, -2.0, 0], [0, 2.0, 0]])) - 1.0) < 1e-12
