"""Launch windows (first CTA start .. last CTA end) of the kernels of two eager forward passes, and the gaps between them."""
import sys, torch
import numpy as np
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
eps = torch.randn(100, NB, L, 3, device="cuda")
use_graph = len(sys.argv) > 1 and sys.argv[1] == "graph"
for _ in range(2): pl.sample(x.clone(), eps, use_graph=False)
torch.cuda.synchronize()
pl.buffer("tc_trace")          # tracing on before the graph is captured (the pointer is baked into the launches)
torch.cuda.synchronize()
xx = x.clone()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); pl.sample(xx, eps, use_graph=use_graph); b.record(); torch.cuda.synchronize()
print("sample ms", a.elapsed_time(b))
tr = pl.buffer("tc_trace").cpu().numpy().astype(np.uint64)
w = tr[1024:1024 + 4096].reshape(-1, 2)
st = (~w[:, 0]).astype(np.int64); en = w[:, 1].astype(np.int64)
ok = w[:, 1] != 0
st, en = st[ok], en[ok]
o = np.argsort(st); st, en = st[o], en[o]
t0 = st[0]
gaps = (st[1:] - en[:-1]) / 1e3
dur = (en - st) / 1e3
print("launch windows:", len(st), " span us", (en[-1] - t0) / 1e3, " sum of windows us", dur.sum(), " sum of gaps us", gaps.sum())
for i in range(800, 840):
    print(f"{(st[i] - t0) / 1e3:9.2f} .. {(en[i] - t0) / 1e3:9.2f}   dur {dur[i]:6.2f}   gap before {0 if i == 0 else gaps[i - 1]:6.2f}")
# per-step summary from end-to-end deltas
nps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ends = en[50 * nps:50 * nps + nps + 1]
print("end deltas:", np.diff(ends / 1e3).round(1).tolist(), " step us", (ends[-1] - ends[0]) / 1e3)
