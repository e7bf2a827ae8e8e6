"""The fp32-storage tensor-core tier (codlad_b200/tf32_tier.py: tcgen05.mma kind::tf32 GEMMs, everything else fp32) against the
unmodified reference's outputs at the configs[1] shape (tests/golden, made by oracle/make_goldens.py with fp32 CPU arithmetic).

TF32 keeps 10 mantissa bits of each GEMM operand -- the same as fp16 -- and the tensor core truncates rather than rounds, so this
tier is NOT more accurate than the f16 tier (measured 7.4e-4 relative vs 4.8e-4): the bars below are the f16 tier's."""
import pytest
import torch

from tests import parity_utils as P
from codlad_b200 import synthetic, weights
from codlad_b200.diffusion import create_diffusion

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tier():
    from codlad_b200.tf32_tier import Tf32Denoiser
    g = P.golden("denoiser_c2_L300x10")
    c = P.members_case(g["meta"])
    den = Tf32Denoiser(weights.init_denoiser_state(0), c["k_neighbors"])
    batch = {k: v.cuda() for k, v in synthetic.collate(c["prot"], [0]).items()}
    den.set_frames(batch, c["members"])
    return den, c, g


def test_tf32_forward_vs_reference(tier):
    den, c, g = tier
    out = den.forward(c["x"].cuda(), c["t"].float().cuda()).cpu()
    ref = torch.from_numpy(g["out"])
    rel, mx = P.rel_err(out, ref), float((out - ref).abs().max())
    print(f"tf32 tier forward: rel {rel:.3e} max-abs {mx:.3e}")
    assert rel < 2e-3 and mx < 1e-2                      # measured 7.4e-4 / 3.6e-3


def test_tf32_sampler_vs_reference_and_graph_replay(tier):
    den, c, _ = tier
    g = P.golden("sampler_c2_L300x10_5")
    L, members, prot_seed, z_seed, noise_seed, steps = (int(v) for v in g["meta"])
    diff = create_diffusion(str(steps))
    z0 = synthetic.latent_noise((members, L, 3), z_seed).cuda()
    nz = synthetic.latent_noise((steps, members, L, 3), noise_seed).cuda().contiguous()
    x = den.sample(diff, z0.clone(), nz, use_graph=False)
    err = P.rel_err(x.cpu(), g["sample_0"])
    print(f"tf32 tier 5-step sampler: rel {err:.3e}")
    assert err < 2e-3                                   # measured 6.6e-4
    xg = den.sample(diff, z0.clone(), nz, use_graph=True)
    assert torch.equal(xg, x)
