import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine, weights
from codlad_b200.diffusion import create_diffusion
torch.set_grad_enabled(False)
L, NB = 300, 10
prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
sd = weights.init_denoiser_state(0)
den = engine.DenoiserEngine(sd, 64)
prot = synthetic.make_protein(L, 1, seed=1002)
pl = engine.Plan(den, 1, NB, L, prec)
pl.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([L]), prot.restype_full[1:-1][None].int(), torch.zeros(NB, dtype=torch.int32))
diff = create_diffusion("100")
pl.set_schedule(diff.timestep_map, diff.coef_table())
x = synthetic.latent_noise((NB, L, 3), 5).cuda()
pl.forward(x, torch.full((NB,), 500.0).cuda())
pl.set_schedule(diff.timestep_map, diff.coef_table())
torch.cuda.synchronize()
for rep in range(3):
    for mode in (0, 1, 2):
        pl.run_edge_kernel(mode, 1)
torch.cuda.synchronize()
print("done")
