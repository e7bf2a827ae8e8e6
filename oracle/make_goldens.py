"""TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/*.npz by running the UNMODIFIED reference
modules (imported from the read-only checkout through oracle/ref_stubs.py) on seeded synthetic
inputs.  The reference ships no tests or golden vectors of its own, so
these fixtures are what pins the oracle (oracle/restate.py) and, through it, the CUDA path.

    python -m oracle.make_goldens            # needs /root/reference; not runnable on the GPU box

Inputs are never stored: they are re-derived from the seeds with codlad_b200.synthetic /
codlad_b200.weights, which is also what the tests do.  Only reference OUTPUTS are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_stubs  # noqa: E402

ref_stubs.install()

from codlad_b200 import synthetic, weights  # noqa: E402
from diffusion_and_flow import create_diffusion  # noqa: E402
import diffusion_and_flow.gaussian_diffusion as gd  # noqa: E402
from models.latent_model import MPNN_models  # noqa: E402
from models.vae_model import VAE, IC_Decoder, IC_Decoder_angle  # noqa: E402
from utils.utils_ic import ic_to_xyz  # noqa: E402
from utils.vq_module import VectorQuantizerEMA, build_quantize  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
C2_CKPT = os.path.join(ref_stubs.REFERENCE_ROOT, "results", "Vae_m1_12-23-23_12345", "model.pt")


def ref_denoiser(k_neighbors=64, seed=0):
    sd = weights.init_denoiser_state(seed)
    m = MPNN_models["mpnn_diffusion"](input_size=3, unconditional=True, diffusion="diffusion", k_neighbors=k_neighbors).eval()
    m.load_state_dict(sd, strict=True)        # also proves the key/shape inventory in weights.py
    return m, sd


def golden_denoiser(name, L, frames, k_neighbors, prot_seed, x_seed, t_list, lengths=None, slim=False):
    """One denoiser forward.  `lengths` (ragged case) builds a batch of proteins of different
    length, padded by the reference's own reshape_and_create_mask."""
    m, _ = ref_denoiser(k_neighbors)
    if lengths is None:
        prot = synthetic.make_protein(L, frames, seed=prot_seed)
        batch = synthetic.collate(prot)
        B, Lmax = frames, L
        mask = torch.ones(B, Lmax, dtype=torch.bool)
    else:
        parts = [synthetic.collate(synthetic.make_protein(n, 1, seed=prot_seed + i)) for i, n in enumerate(lengths)]
        batch = {"CG_nxyz": torch.cat([p["CG_nxyz"] for p in parts]), "num_CGs": torch.tensor(lengths),
                 "CG_nbr_list": parts[0]["CG_nbr_list"]}
        B, Lmax = len(lengths), max(lengths)
        mask = torch.arange(Lmax)[None, :] < torch.tensor(lengths)[:, None]
    x = synthetic.latent_noise((B, Lmax, 3), x_seed)
    batch["randn"] = torch.zeros(B, Lmax)
    t = torch.tensor(t_list, dtype=torch.int64)
    with torch.no_grad():
        X = torch.zeros(B, Lmax, 3)
        off = 0
        for b in range(B):
            n = int(batch["num_CGs"][b])
            X[b, :n] = batch["CG_nxyz"][off:off + n, 1:]
            off += n
        E, E_idx = m.features(X, mask.int(), torch.arange(Lmax)[None].expand(B, -1), torch.ones(B, Lmax))
        D_nb, _, _ = m.features._dist(X, mask.int())
        out = m(x, t, None, mask=mask, batch=batch)
    if slim:      # large shapes: every 8th distance row, no feature sample (the small cases pin those stages)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), E_idx=E_idx.numpy().astype(np.int16), D_nb=D_nb.numpy()[:, ::8], out=out.numpy(),
                            meta=np.array([L, frames, k_neighbors, prot_seed, x_seed] + list(t_list)))
        print(name, "out", tuple(out.shape), float(out.abs().max()))
        return
    np.savez_compressed(os.path.join(OUT, name + ".npz"), E_idx=E_idx.numpy().astype(np.int16), D_nb=D_nb.numpy(),
                        E_sample=E[:, ::7, ::5].numpy(), out=out.numpy(),
                        meta=np.array([L, frames, k_neighbors, prot_seed, x_seed] + list(t_list)))
    print(name, "out", tuple(out.shape), float(out.abs().max()))


def golden_sampler(name, L, prot_seed, z_seed, noise_seed, steps=100):
    """Full p_sample_loop through SpacedDiffusion with the per-step noise injected."""
    m, _ = ref_denoiser()
    prot = synthetic.make_protein(L, 1, seed=prot_seed)
    batch = synthetic.collate(prot)
    batch["randn"] = torch.zeros(1, L)
    mask = torch.ones(1, L, dtype=torch.bool)
    z = synthetic.latent_noise((1, L, 3), z_seed)
    noises = synthetic.latent_noise((steps, 1, L, 3), noise_seed)
    diffusion = create_diffusion(str(steps))
    order = list(range(steps))[::-1]
    it = iter(order)
    real = gd.th.randn_like
    gd.th.randn_like = lambda x: noises[next(it)].to(x)      # reference draws at gaussian_diffusion.py:440
    keep = {}
    try:
        with torch.no_grad():
            for s, out in zip(order, diffusion.p_sample_loop_progressive(
                    m.forward, z.shape, z, clip_denoised=False,
                    model_kwargs=dict(y=None, mask=mask, batch=batch), device="cpu")):
                if s in (steps - 1, steps // 2, 1, 0):
                    keep[f"sample_{s}"] = out["sample"].numpy()
    finally:
        gd.th.randn_like = real
    np.savez_compressed(os.path.join(OUT, name + ".npz"), timestep_map=np.array(diffusion.timestep_map),
                        meta=np.array([L, prot_seed, z_seed, noise_seed, steps]), **keep)
    print(name, {k: float(np.abs(v).max()) for k, v in keep.items()})


def golden_denoiser_members(name, L, members, k_neighbors, prot_seed, x_seed, t_list):
    """One denoiser forward on ONE frame with `members` batch rows (ensemble members: the reference sees `members` copies of
    the frame in its batch dict) -- the BASELINE configs[1] / configs[3] geometry.  Only the model output and the
    neighbour indices of the frame are stored (the distances are pinned bit-exactly against the oracle elsewhere)."""
    m, _ = ref_denoiser(k_neighbors)
    prot = synthetic.make_protein(L, 1, seed=prot_seed)
    batch = synthetic.collate(prot, frames=[0] * members)
    mask = torch.ones(members, L, dtype=torch.bool)
    x = synthetic.latent_noise((members, L, 3), x_seed)
    batch["randn"] = torch.zeros(members, L)
    t = torch.tensor(t_list, dtype=torch.int64)
    with torch.no_grad():
        X = prot.ca_full[:, 1:-1].contiguous()
        _, E_idx = m.features(X, torch.ones(1, L, dtype=torch.int32), torch.arange(L)[None], torch.ones(1, L))
        D_nb, _, _ = m.features._dist(X, torch.ones(1, L, dtype=torch.int32))
        out = m(x, t, None, mask=mask, batch=batch)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), E_idx=E_idx.numpy().astype(np.int16), D_nb=D_nb.numpy().astype(np.float32)[:, ::8],
                        out=out.numpy(), meta=np.array([L, members, k_neighbors, prot_seed, x_seed] + list(t_list)))
    print(name, "out", tuple(out.shape), float(out.abs().max()))


def golden_sampler_members(name, L, members, prot_seed, z_seed, noise_seed, steps):
    """`steps`-step p_sample_loop of `members` ensemble members on one frame (configs[1] geometry, short schedule)."""
    m, _ = ref_denoiser()
    prot = synthetic.make_protein(L, 1, seed=prot_seed)
    batch = synthetic.collate(prot, frames=[0] * members)
    batch["randn"] = torch.zeros(members, L)
    mask = torch.ones(members, L, dtype=torch.bool)
    z = synthetic.latent_noise((members, L, 3), z_seed)
    noises = synthetic.latent_noise((steps, members, L, 3), noise_seed)
    diffusion = create_diffusion(str(steps))
    it = iter(list(range(steps))[::-1])
    real = gd.th.randn_like
    gd.th.randn_like = lambda x: noises[next(it)].to(x)
    try:
        with torch.no_grad():
            sample = diffusion.p_sample_loop(m.forward, z.shape, z, clip_denoised=False,
                                             model_kwargs=dict(y=None, mask=mask, batch=batch), device="cpu")
    finally:
        gd.th.randn_like = real
    np.savez_compressed(os.path.join(OUT, name + ".npz"), timestep_map=np.array(diffusion.timestep_map), sample_0=sample.numpy(),
                        meta=np.array([L, members, prot_seed, z_seed, noise_seed, steps]))
    print(name, float(sample.abs().max()))


def export_c2_decoder():
    ck = torch.load(C2_CKPT, map_location="cpu")
    want = weights.ic_decoder_shapes(False)
    out = {}
    for k, shape in want.items():
        assert tuple(ck[k].shape) == tuple(shape), k
        out[k] = ck[k].float().numpy()
    np.savez_compressed(os.path.join(OUT, "ic_decoder_c2.npz"), **out)
    print("ic_decoder_c2", len(out), sum(v.size for v in out.values()))


def decode_state(angle_variant, use_c2=False):
    """Random-init (gain 0.5, angles O(1): coordinates are well conditioned) or, with use_c2, the
    shipped C2 checkpoint's `equivaraintconv.*` tensors loaded into IC_Decoder (real weight
    magnitudes; they drive angles to O(1e6) rad on synthetic inputs, so only ic_recon is compared,
    with a relative tolerance)."""
    sd = weights.init_vae_decode_state(0, angle_variant=angle_variant)
    if use_c2:
        c2 = np.load(os.path.join(OUT, "ic_decoder_c2.npz"))
        for k in c2.files:
            sd[k] = torch.from_numpy(c2[k])
    return sd


def golden_decode(name, L, frames, prot_seed, lat_seed, angle_variant, use_c2=False):
    sd = decode_state(angle_variant, use_c2)
    dec = (IC_Decoder_angle if angle_variant else IC_Decoder)(n_atom_basis=36, n_rbf=15, cutoff=21.0, num_conv=4, activation="swish")
    vae = VAE(5, 36, encoder=None, quantize=build_quantize("vqvae", 4096, 3, 0.25, 0.99), equivaraintconv=dec, vqdim=3).eval()
    res = vae.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"map_in.weight", "map_in.bias"}, res
    prot = synthetic.make_protein(L, frames, seed=prot_seed)
    batch = synthetic.collate(prot)
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    latent = mean + std * synthetic.latent_noise((frames, L, 3), lat_seed)
    mask = torch.ones(frames, L, dtype=torch.bool)
    with torch.no_grad():
        _, ic_recon = vae.latent_decode(latent, mask, batch)
        xyz = ic_to_xyz(batch["OG_CG_nxyz"].reshape(-1, L + 2, 4), ic_recon.reshape(-1, L, 13, 3), prot.info)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ic_recon=ic_recon.numpy(), xyz=xyz.numpy(),
                        meta=np.array([L, frames, prot_seed, lat_seed, int(angle_variant)]))
    print(name, "ic", tuple(ic_recon.shape), "xyz", tuple(xyz.shape), float(ic_recon[..., 1:].abs().max()))


def golden_vq(name, n, seed):
    """Index parity of the restated third-party VQ against the in-repo VectorQuantizerEMA
    (utils/vq_module.py:56-71), the only reference-side implementation that exists offline."""
    sd = weights.init_vae_decode_state(0)
    cb = sd["quantize._codebook.embed"][0]
    vq = VectorQuantizerEMA(4096, 3, 0.25, 0.99).eval()
    vq.embeddings.copy_(cb)
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    x = mean + std * synthetic.latent_noise((n, 3), seed)
    with torch.no_grad():
        _, idx, _ = vq(x[None])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), idx=idx.numpy().astype(np.int16), meta=np.array([n, seed]))
    print(name, idx.shape)


def golden_ic_large_angle(name, L, seed):
    """ic_to_xyz alone with angles of magnitude up to 1e5 rad (random-init decoders produce such
    values, SURVEY 'hard parts'): full-range sin/cos reduction must agree."""
    prot = synthetic.make_protein(L, 2, seed=seed)
    g = torch.Generator().manual_seed(seed)
    ic = torch.randn(2, L, 13, 3, generator=g)
    ic[..., 0] = 1.0 + 0.5 * torch.rand(2, L, 13, generator=g)
    ic[..., 1:] *= torch.tensor([1.0, 1e2, 1e5])[torch.randint(0, 3, (2, L, 13, 1), generator=g)].expand(-1, -1, -1, 2)
    batch = synthetic.collate(prot)
    with torch.no_grad():
        xyz = ic_to_xyz(batch["OG_CG_nxyz"].reshape(-1, L + 2, 4), ic, prot.info)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), ic=ic.numpy(), xyz=xyz.numpy(), meta=np.array([L, seed]))
    print(name, tuple(xyz.shape))


def golden_training_losses(name, B, L, seed):
    """training_losses of the unmodified reference (gaussian_diffusion.py:549-725, MSE loss + learned-range VB term) for a fixed
    "model output": pins q_sample, the posterior, normal_kl, the discretised likelihood and the masked means.  Inputs are
    re-derived from the seed by the test (same generator calls)."""
    diffusion = create_diffusion(timestep_respacing="")                  # training uses the unspaced 1000-step chain
    g = torch.Generator().manual_seed(seed)
    x_start = torch.randn(B, L, 3, generator=g)
    noise = torch.randn(B, L, 3, generator=g)
    model_out = 0.7 * torch.randn(B, L, 6, generator=g)
    t = torch.tensor([0, 1, 500, 999, 37, 0][:B])
    mask = torch.ones(B, L, dtype=torch.bool)
    mask[1, L - 5:] = False
    mask[-1, L // 2:] = False
    model = lambda x_t, tt, **kw: model_out
    with torch.no_grad():
        terms = diffusion.training_losses(model, x_start, t, dict(mask=mask), noise=noise)
        terms_nomask = diffusion.training_losses(model, x_start, t, dict(), noise=noise)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), mse=terms["mse"].numpy(), vb=terms["vb"].numpy(), loss=terms["loss"].numpy(),
                        loss_nomask=terms_nomask["loss"].numpy(), meta=np.array([B, L, seed]))
    print(name, terms["loss"].tolist())


def synthetic_all_atom(L, frames, seed, sigma):
    """Reference / generated all-atom structures for the metric goldens: the reference structure is the synthetic protein rebuilt
    from seeded internal coordinates by the reference's own ic_to_xyz; the generated one is that structure plus N(0, sigma) noise
    (bonds break as sigma grows).  Atomic numbers: N, CA, C, O backbone, carbon / sulphur side chains by slot parity."""
    prot = synthetic.make_protein(L, frames, seed=seed)
    batch = synthetic.collate(prot)
    g = torch.Generator().manual_seed(seed + 1)
    ic = torch.stack([1.2 + 0.4 * torch.rand(frames * L, 13, generator=g), 1.8 + 0.5 * torch.rand(frames * L, 13, generator=g),
                      6.28 * torch.rand(frames * L, 13, generator=g) - 3.14], -1)
    with torch.no_grad():
        xyz = ic_to_xyz(batch["OG_CG_nxyz"].reshape(-1, L + 2, 4), ic.reshape(frames, L, 13, 3), prot.info)     # [frames, Na, 3]
    na = xyz.shape[1]
    z = torch.tensor([7, 6, 6, 8] + [6, 16, 1, 6, 8, 7, 1, 6, 1, 6], dtype=torch.int64).repeat((na + 13) // 14)[:na]
    gen = xyz + sigma * torch.randn(xyz.shape, generator=g)
    return xyz.reshape(-1, 3), gen.reshape(-1, 3), z.repeat(frames), [na] * frames


def golden_sample_qualities(name, L, frames, seed, sigma):
    """eval_sample_qualities of the unmodified reference (utils/protein_module.py:335-364), one reconstruction per structure as
    test.py:178-181 calls it.  ASE is absent here: a 3-method stand-in for ase.Atoms is injected into the reference module."""
    import utils.protein_module as pm

    class Atoms:
        def __init__(self, numbers, positions):
            self.z, self.x = np.asarray(numbers), np.asarray(positions, dtype=np.float64)
        def get_positions(self): return self.x
        def get_atomic_numbers(self): return self.z
        def __len__(self): return len(self.z)

    pm.Atoms = Atoms
    ref, gen, z, num = synthetic_all_atom(L, frames, seed, sigma)
    out, o = [], 0
    for na in num:
        ra = Atoms(z[o:o + na].numpy(), ref[o:o + na].numpy())
        ga = Atoms(z[o:o + na].numpy(), gen[o:o + na].numpy())
        all_rmsds, heavy_rmsds, vr, var, gvr, gavr = pm.eval_sample_qualities(ra, [ga])
        rm = pm.compute_rmsd([ga], ra, [0])[0]
        out.append([vr, var, gvr[0], gavr[0], rm[0], rm[1]])
        o += na
    # (the structures come out of the reference's ic_to_xyz, so -- unlike the other goldens -- the inputs are stored too)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), q=np.array(out, dtype=np.float64), meta=np.array([L, frames, seed]), sigma=np.array(sigma),
                        xyz_ref=ref.numpy(), xyz_gen=gen.numpy(), z=z.numpy().astype(np.int16), num_atoms=np.array(num))
    print(name, np.array(out).round(4).tolist())


def golden_eval_losses(name, seed):
    """inter_result / clash_result / ged_result of the unmodified reference (test.py:97-146).  test.py itself cannot be imported here
    (mdtraj, ase, torchdiffeq are absent), so the three function definitions are taken from its syntax tree and executed as they are."""
    import ast
    src = open(os.path.join(ref_stubs.REFERENCE_ROOT, "test.py")).read()
    want = ("inter_result", "clash_result", "ged_result")
    mod = ast.Module(body=[n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in want], type_ignores=[])
    ns = {"torch": torch, "EPS": 1e-7}
    exec(compile(mod, "reference/test.py", "exec"), ns)
    c = synthetic.eval_loss_case(seed)
    li, lp = ns["inter_result"](c["inter"], c["pipi"], c["recon"])
    li0, _ = ns["inter_result"](c["inter"], c["pipi"][:0], c["recon"])
    clash = ns["clash_result"](c["edge"], c["nbr"], c["recon"], c["bb"])
    ged = ns["ged_result"](c["recon"], c["xyz"], c["edge"])
    u, cnt = torch.cat((c["edge"], c["nbr"])).unique(dim=0, return_counts=True)
    vals = np.array([float(li), float(lp), float(li0), float(clash), float(ged), float((cnt == 1).sum())], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), vals=vals, meta=np.array([seed]))
    print(name, vals.tolist())


GOLDENS = {
    "eval_losses": lambda n: golden_eval_losses(n, 8101),
    "denoiser_L64_B1": lambda n: golden_denoiser(n, 64, 1, 64, 1001, 2001, [717]),
    "denoiser_L70_B2": lambda n: golden_denoiser(n, 70, 2, 64, 1002, 2002, [999, 10]),
    "denoiser_L100_K48": lambda n: golden_denoiser(n, 100, 1, 48, 1004, 2004, [505]),
    "denoiser_L40_short": lambda n: golden_denoiser(n, 40, 2, 64, 1006, 2006, [0, 303]),           # K = L < 64
    "denoiser_ragged": lambda n: golden_denoiser(n, 80, 0, 64, 1005, 2005, [61, 989], lengths=[80, 70]),
    "sampler_L64_100": lambda n: golden_sampler(n, 64, 1001, 2001, 3001, 100),
    "decode_L64_N6": lambda n: golden_decode(n, 64, 2, 1001, 4001, False),
    "decode_L64_K4": lambda n: golden_decode(n, 64, 1, 1001, 4002, True),
    "decode_L64_N6_c2": lambda n: golden_decode(n, 64, 1, 1001, 4003, False, use_c2=True),
    "vq_20000": lambda n: golden_vq(n, 20000, 5001),
    "ic_large_angle_L48": lambda n: golden_ic_large_angle(n, 48, 6001),
    "training_losses_B6": lambda n: golden_training_losses(n, 6, 24, 7001),
    "sample_qualities_exact": lambda n: golden_sample_qualities(n, 40, 2, 8001, 0.0),
    "sample_qualities_noisy": lambda n: golden_sample_qualities(n, 40, 3, 8002, 0.08),
    # BASELINE.json shapes: configs[1] (300 residues x 10 members), configs[2] (500-residue proteins), configs[3] (2000 residues, k = 48)
    "denoiser_c2_L300x10": lambda n: golden_denoiser_members(n, 300, 10, 64, 1002, 2102, [999, 989, 747, 500, 500, 252, 131, 10, 0, 0]),
    "sampler_c2_L300x10_5": lambda n: golden_sampler_members(n, 300, 10, 1002, 2103, 3103, 5),
    "denoiser_c3_L500_B3": lambda n: golden_denoiser(n, 500, 3, 64, 1013, 2113, [999, 421, 7], slim=True),
    "denoiser_c4_L2000_K48x2": lambda n: golden_denoiser_members(n, 2000, 2, 48, 1014, 2114, [868, 40]),
    "decode_c2_L300_N6": lambda n: golden_decode(n, 300, 1, 1002, 4102, False),
    "decode_c4_L2000_K4": lambda n: golden_decode(n, 2000, 1, 1014, 4114, True),
}


def main():
    """python -m oracle.make_goldens [name ...]   (no names = every golden)"""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    names = sys.argv[1:] or list(GOLDENS)
    if "decode_L64_N6_c2" in names or not sys.argv[1:]:
        export_c2_decoder()
    for n in names:
        GOLDENS[n](n)


if __name__ == "__main__":
    main()
