import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine, weights
torch.set_grad_enabled(False)
L, NB = 300, 10
sd = weights.init_denoiser_state(0)
den = engine.DenoiserEngine(sd, 64)
prot = synthetic.make_protein(L, 1, seed=1002)
pl = engine.Plan(den, 1, NB, L, "f16")
pl.set_frames(prot.ca_full[:, 1:-1].contiguous(), torch.tensor([L]), prot.restype_full[1:-1][None].int(), torch.zeros(NB, dtype=torch.int32))
x = synthetic.latent_noise((NB, L, 3), 5).cuda()
t = torch.full((NB,), 500.0).cuda()
for _ in range(3): pl.forward(x, t)
torch.cuda.synchronize()
pl.buffer("tc_trace")
torch.cuda.synchronize()
stop = int(sys.argv[1]) if len(sys.argv) > 1 else 3
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); pl.forward_partial(x, t, stop); b.record(); torch.cuda.synchronize()
print("partial forward us", a.elapsed_time(b) * 1e3)
tr = pl.buffer("tc_trace").cpu().tolist()
n = tr[512]
cn = ["control start", "?", "operands ready", "?", "phase committed", "?", "  weight tile landed", "  gemm issued", "  hidden chunk ready"]
prev = None; t0 = None
for v in tr[513:513 + n]:
    v &= (1 << 64) - 1
    tt, ev = v >> 8, v & 0xff
    if ev == 0: t0 = tt; prev = None; print("--- kernel")
    print(f"{(tt - t0) / 1000:8.2f} us  (+{0 if prev is None else (tt - prev) / 1000:6.2f})  {cn[ev]}")
    prev = tt
