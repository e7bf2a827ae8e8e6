"""Needs a trace build: bash tools/dev/build_variant.sh trace -DCB2_TRACE_EPI; CB2_LIB=codlad_b200/_variants/lib_trace.so"""
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
names = ["E1 begin", "E1 acc ready", "E1 math done", "E1 handed off", "E2 begin", "E2 acc ready", "E2 math done", "E2 handed off", "drain begin", "drain acc ready", "drain done"]
t0 = None
prev = None
for v in tr[1:1 + n]:
    v &= (1 << 64) - 1
    t, code = v >> 8, v & 0xff
    ev, s = code >> 2, code & 3
    if t0 is None: t0 = t
    print(f"{(t - t0) / 1000:8.2f} us  (+{0 if prev is None else (t - prev) / 1000:6.2f})  slot {s}  {names[ev]}")
    prev = t
import numpy as np
w = np.array(tr[600:600 + 2 * 148], dtype=np.int64).reshape(-1, 2)
st, en = w[:, 0], w[:, 1]
print("CTA start spread us", (st.max() - st.min()) / 1e3, " first start -> last end us", (en.max() - st.min()) / 1e3)
d = (en - st) / 1e3
print("CTA durations us: min %.1f mean %.1f max %.1f" % (d.min(), d.mean(), d.max()))
print("end times rel. first start, sorted:", np.sort((en - st.min()) / 1e3).round(1).tolist()[::8])
