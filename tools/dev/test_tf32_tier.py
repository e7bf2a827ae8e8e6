"""Dev check: TF32 inference tier vs the reference goldens at the configs[1] shape + timing of one graph-replayed step."""
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, torch
import parity_utils as P
from codlad_b200 import synthetic, weights
from codlad_b200.diffusion import create_diffusion
from codlad_b200.tf32_tier import Tf32Denoiser

g = P.golden("denoiser_c2_L300x10")
c = P.members_case(g["meta"])
den = Tf32Denoiser(weights.init_denoiser_state(0), c["k_neighbors"])
batch = {k: v.cuda() for k, v in synthetic.collate(c["prot"], [0]).items()}
den.set_frames(batch, c["members"])
out = den.forward(c["x"].cuda(), c["t"].float().cuda()).cpu()
ref = torch.from_numpy(g["out"])
print("forward tf32: rel %.3e max-abs %.3e" % (P.rel_err(out, ref), float((out - ref).abs().max())))

g = P.golden("sampler_c2_L300x10_5")
L, members, prot_seed, z_seed, noise_seed, steps = (int(v) for v in g["meta"])
diff = create_diffusion(str(steps))
z0 = synthetic.latent_noise((members, L, 3), z_seed).cuda()
nz = synthetic.latent_noise((steps, members, L, 3), noise_seed).cuda().contiguous()
x = den.sample(diff, z0.clone(), nz, use_graph=False)
print("sampler tf32 eager: rel %.3e" % P.rel_err(x.cpu(), g["sample_0"]))
xg = den.sample(diff, z0.clone(), nz, use_graph=True)
print("graph == eager:", bool(torch.equal(xg, x)), "rel %.3e" % P.rel_err(xg.cpu(), g["sample_0"]))

diff = create_diffusion("100")
nz = torch.randn(100, members, L, 3, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    den.sample(diff, z0.clone(), nz, use_graph=True)
    torch.cuda.synchronize(); dt = time.time() - t0
    print("100 steps: %.1f ms -> %.1f k res/s" % (dt * 1e3, members * L / dt / 1e3))
