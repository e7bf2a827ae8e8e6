"""Shared helpers for parity tests: golden loading, seeded input reconstruction, tie-aware k-NN
comparison (SURVEY.md 'hard parts': torch.topk leaves the order of exactly tied distances
unspecified, so indices are compared position-wise only where the distance row is strictly
increasing, and as sets inside runs of equal distance / across the K boundary)."""
import os

import numpy as np
import torch

from codlad_b200 import synthetic, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def denoiser_case(meta, lengths=None):
    """Rebuild the inputs of oracle/make_goldens.golden_denoiser from its meta row."""
    L, frames, k_nb, prot_seed, x_seed = (int(v) for v in meta[:5])
    t = torch.tensor([int(v) for v in meta[5:]], dtype=torch.int64)
    if lengths is None:
        prot = synthetic.make_protein(L, frames, seed=prot_seed)
        X = prot.ca_full[:, 1:-1].contiguous()
        z = prot.restype_full[1:-1][None].expand(frames, -1).contiguous()
        mask = torch.ones(frames, L, dtype=torch.bool)
    else:
        Lmax = max(lengths)
        X = torch.zeros(len(lengths), Lmax, 3)
        z = torch.zeros(len(lengths), Lmax, dtype=torch.int64)
        for b, n in enumerate(lengths):
            p = synthetic.make_protein(n, 1, seed=prot_seed + b)
            X[b, :n] = p.ca_full[0, 1:-1]
            z[b, :n] = p.restype_full[1:-1]
        mask = torch.arange(Lmax)[None, :] < torch.tensor(lengths)[:, None]
    x = synthetic.latent_noise(tuple(mask.shape) + (3,), x_seed)
    return dict(X=X, cg_z=z, mask=mask, x=x, t=t, k_neighbors=k_nb)


def members_case(meta):
    """Rebuild the inputs of oracle/make_goldens.golden_denoiser_members: ONE frame, `members` batch rows on it."""
    L, members, k_nb, prot_seed, x_seed = (int(v) for v in meta[:5])
    t = torch.tensor([int(v) for v in meta[5:]], dtype=torch.int64)
    prot = synthetic.make_protein(L, 1, seed=prot_seed)
    X = prot.ca_full[:, 1:-1].contiguous()                      # [1, L, 3]
    z = prot.restype_full[1:-1][None].contiguous()              # [1, L]
    x = synthetic.latent_noise((members, L, 3), x_seed)
    return dict(prot=prot, X=X, cg_z=z, x=x, t=t, k_neighbors=k_nb, members=members, L=L)


def knn_tie_aware_equal(idx_a, idx_b, d_ref, valid_rows=None):
    """idx_* [R,K] integer arrays, d_ref [R,K] the reference's sorted distances for those rows."""
    idx_a, idx_b, d_ref = (np.asarray(v) for v in (idx_a, idx_b, d_ref))
    bad = 0
    for r in range(idx_a.shape[0]):
        if valid_rows is not None and not valid_rows[r]:
            continue
        if np.array_equal(idx_a[r], idx_b[r]):
            continue
        d = d_ref[r]
        K = d.shape[0]
        k = 0
        while k < K:
            e = k
            while e + 1 < K and d[e + 1] == d[k]:
                e += 1
            seg_a, seg_b = set(idx_a[r, k:e + 1].tolist()), set(idx_b[r, k:e + 1].tolist())
            if seg_a != seg_b and e != K - 1:      # a tie run touching the K boundary may pick any tied candidate
                bad += 1
                break
            k = e + 1
    return bad == 0


def rel_err(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rmsd(a, b):
    a, b = torch.as_tensor(a, dtype=torch.float64), torch.as_tensor(b, dtype=torch.float64)
    return float(((a - b) ** 2).sum(-1).mean().sqrt())


def decode_state(angle_variant, use_c2=False):
    sd = weights.init_vae_decode_state(0, angle_variant=angle_variant)
    if use_c2:
        c2 = golden("ic_decoder_c2")
        for k in c2.files:
            sd[k] = torch.from_numpy(c2[k])
    return sd
