"""Per-warpgroup timeline of CTA 0 of one per-edge kernel.  Needs a trace build:
   bash tools/dev/build_variant.sh tracewg -DCB2_WITH_WG -DCB2_TRACE_WG ; CB2_EDGE_WG=1 CB2_LIB=codlad_b200/_variants/lib_tracewg.so python tools/dev/trace_wg.py MODE"""
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
names = ["tile start", "load landed, MMA1 issued", "acc1 ready", "E1 closed", "acc2 ready", "E2 closed", "acc3 ready", "tile closed", "E3 pass A done"]
t0 = min((tr[g * 256 + 1] & ((1 << 64) - 1)) >> 8 for g in range(4) if tr[g * 256] > 0)
for g in range(4):
    n = tr[g * 256]
    prev = None
    print(f"--- warpgroup {g}: {n} events")
    for v in tr[g * 256 + 1:g * 256 + 1 + n]:
        v &= (1 << 64) - 1
        t, ev = v >> 8, v & 0xff
        print(f"{(t - t0) / 1000:8.2f} us  (+{0 if prev is None else (t - prev) / 1000:6.2f})  {names[ev]}")
        prev = t
