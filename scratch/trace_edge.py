import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine, weights
from codlad_b200.diffusion import create_diffusion
torch.set_grad_enabled(False)
L, NB = 300, 10
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
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
for _ in range(3): pl.run_edge_kernel(mode, 1)
torch.cuda.synchronize()
pl.buffer("tc_trace")          # allocates + zeroes; tracing on from now
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); pl.run_edge_kernel(mode, 1); b.record(); torch.cuda.synchronize()
print("kernel us", a.elapsed_time(b) * 1e3)
tr = pl.buffer("tc_trace").cpu().tolist()
n = tr[0]
names = ["E1 loads issued", "E1 acc ready", "E2 loads issued", "E2 acc ready", "E3 begin", "E3 acc ready", "stage done"]
t0 = None
prev = None
for v in tr[1:1 + n]:
    v &= (1 << 64) - 1
    t, code = v >> 8, v & 0xff
    ev, s = code >> 2, code & 3
    if t0 is None: t0 = t
    print(f"{(t - t0) / 1000:8.2f} us  (+{0 if prev is None else (t - prev) / 1000:6.2f})  slot {s}  {names[ev]}")
    prev = t

print("---- control thread")
n = tr[512]
cn = ["S1 wait load", "S1 loaded", "S1 mma issued", "E1 wait", "E1 done seen", "MMA2 issued", "E2 wait", "E2 done seen", "RED issued", "RED complete", "next load issued", "E3 wait", "E3 done seen"]
prev = None
for v in tr[513:513 + n]:
    v &= (1 << 64) - 1
    t, code = v >> 8, v & 0xff
    ev, s = code >> 2, code & 3
    print(f"{(t - t0) / 1000:8.2f} us  (+{0 if prev is None else (t - prev) / 1000:6.2f})  slot {s}  {cn[ev]}")
    prev = t
