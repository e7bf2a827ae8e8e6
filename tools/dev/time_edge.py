"""Warm (back-to-back) time of the three per-edge kernels at configs[1]; CB2_LIB selects an ablation build."""
import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine, weights
from codlad_b200.diffusion import create_diffusion
torch.set_grad_enabled(False)
L, NB = 300, 10
sd = weights.init_denoiser_state(0)
den = engine.DenoiserEngine(sd, 64)
prot = synthetic.make_protein(L, 1, seed=1002)
pl = engine.Plan(den, 1, NB, L, "f16")
pl.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([L]), prot.restype_full[1:-1][None].int(), torch.zeros(NB, dtype=torch.int32))
diff = create_diffusion("100")
pl.set_schedule(diff.timestep_map, diff.coef_table())
x = synthetic.latent_noise((NB, L, 3), 5).cuda()
pl.forward(x, torch.full((NB,), 500.0).cuda())
pl.set_schedule(diff.timestep_map, diff.coef_table())
out = []
for mode in (0, 1, 2):
    for _ in range(5): pl.run_edge_kernel(mode, 1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): pl.run_edge_kernel(mode, 1)
    b.record(); torch.cuda.synchronize()
    out.append(a.elapsed_time(b) / 50 * 1e3)
print(sys.argv[1] if len(sys.argv) > 1 else "", "us per launch (warm): msg_enc %.1f  edge_update %.1f  msg_dec %.1f" % tuple(out))
