import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine, weights
from codlad_b200.diffusion import create_diffusion
torch.set_grad_enabled(False)
L, NB = 300, 10
den = engine.DenoiserEngine(weights.init_denoiser_state(0), 64)
prot = synthetic.make_protein(L, 1, seed=1002)
pl = engine.Plan(den, 1, NB, L, "f16")
pl.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([L]), prot.restype_full[1:-1][None].int(), torch.zeros(NB, dtype=torch.int32))
diff = create_diffusion("100")
pl.set_schedule(diff.timestep_map, diff.coef_table())
x = synthetic.latent_noise((NB, L, 3), 5).cuda()
t = torch.full((NB,), 500.0).cuda()
pl.forward(x, t); pl.set_schedule(diff.timestep_map, diff.coef_table())
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps
for mode in (0, 1, 2):
    print(f"edge mode {mode}: {timeit(lambda: pl.run_edge_kernel(mode, 1)):.1f} us (back-to-back, warm L2)")
prev = 0.0
names = ["init"] + sum([[f"enc{l}.msg", f"enc{l}.node", f"enc{l}.edge"] for l in range(3)], []) + sum([[f"dec{l}.msg", f"dec{l}.node"] for l in range(3)], [])
for stop in range(1, 17):
    tt = timeit(lambda: pl.forward_partial(x, t, stop), 20)
    print(f"{names[stop-1]:10s} cumulative {tt:7.1f} us  delta {tt - prev:6.1f}")
    prev = tt
# whole sampling loop via graph
noise = torch.randn(100, NB, L, 3, device="cuda")
xx = x.clone()
pl.sample(xx, noise, True); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); pl.sample(xx, noise, True); b.record(); torch.cuda.synchronize()
print("graph 100 steps:", a.elapsed_time(b), "ms")
