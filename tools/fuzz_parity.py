"""Randomised geometry sweep of the denoiser path: for random (frames, ragged lengths, members, k) the f16 tier must track the fp32 tier
(which the parity tests pin to the oracle), outputs must be finite, padded rows must be exactly what the masked path defines, and a
3-step sampling loop must replay bit-identically from its CUDA graph.  Usage: python tools/fuzz_parity.py [n_cases] [seed]"""
import random
import sys

import torch

sys.path.insert(0, ".")
from codlad_b200 import engine as eng, synthetic, weights          # noqa: E402
from codlad_b200.diffusion import create_diffusion                  # noqa: E402

torch.set_grad_enabled(False)
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
sd = weights.init_denoiser_state(0)
engines = {}
worst = 0.0
for case in range(n_cases):
    k = rnd.choice([16, 30, 48, 64])
    F = rnd.choice([1, 1, 2, 3, 4])
    Lmax = rnd.choice([rnd.randint(2, 40), rnd.randint(41, 200), rnd.randint(201, 700)])
    lengths = [Lmax] + [rnd.randint(max(1, Lmax // 2), Lmax) for _ in range(F - 1)]
    members = rnd.randint(1, 3) if F > 1 else rnd.randint(1, 12)
    NB = F * members
    X = torch.zeros(F, Lmax, 3)
    z = torch.zeros(F, Lmax, dtype=torch.int32)
    for f, n in enumerate(lengths):
        p = synthetic.make_protein(n, 1, seed=7000 + 13 * case + f, compact=rnd.choice([0.0, 0.002]))
        X[f, :n] = p.ca_full[0, 1:-1]
        z[f, :n] = p.restype_full[1:-1].int()
    frame_of = torch.arange(F, dtype=torch.int32).repeat(members)
    x = synthetic.latent_noise((NB, Lmax, 3), 100 + case).cuda()
    t = torch.tensor([float(rnd.randint(0, 999)) for _ in range(NB)]).cuda()
    den = engines.setdefault(k, eng.DenoiserEngine(sd, k))
    outs = {}
    for prec in ("fp32", "f16"):
        plan = eng.Plan(den, F, NB, Lmax, prec)
        plan.set_frames(X, torch.tensor(lengths), z, frame_of)
        outs[prec] = plan.forward(x, t).cpu()
        if prec == "f16":
            diff = create_diffusion("3")
            plan.set_schedule(diff.timestep_map, diff.coef_table())
            nz = synthetic.latent_noise((3, NB, Lmax, 3), 500 + case).cuda()
            a = plan.sample(x.clone(), nz, use_graph=False).cpu()
            b = plan.sample(x.clone(), nz, use_graph=True).cpu()
            assert torch.equal(a, b), f"case {case}: graph replay differs"
        del plan
    valid = (torch.arange(Lmax)[None, :] < torch.tensor(lengths)[:, None])[frame_of.long()]
    assert torch.isfinite(outs["f16"]).all() and torch.isfinite(outs["fp32"]).all(), f"case {case}: non-finite output"
    # rows whose frame has fewer than k residues see the reference's D_max padding: compared like all others
    a, b = outs["f16"][valid].double(), outs["fp32"][valid].double()
    rel = float((a - b).norm() / b.norm().clamp_min(1e-30))
    worst = max(worst, rel)
    print(f"case {case:3d}: F={F} lengths={lengths} members={members} k={k}  f16 vs fp32 rel {rel:.2e}")
    assert rel < 4e-3, f"case {case}: f16 tier off by {rel}"
print(f"{n_cases} cases, worst f16-vs-fp32 relative difference {worst:.2e}")
