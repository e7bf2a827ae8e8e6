import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, engine, weights
torch.set_grad_enabled(False)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 2
KN = int(sys.argv[3]) if len(sys.argv) > 3 else 64
sd = weights.init_denoiser_state(0)
den = engine.DenoiserEngine(sd, KN)
prot = synthetic.make_protein(L, 1, seed=77)
X = prot.ca_full[:, 1:-1].contiguous()
z = prot.restype_full[1:-1][None]
x = synthetic.latent_noise((NB, L, 3), 5)
t = torch.linspace(3, 999, NB)
plans = {}
for prec in ("fp32", "f16"):
    pl = engine.Plan(den, 1, NB, L, prec)
    pl.set_frames(X, torch.tensor([L]), z.int(), torch.zeros(NB, dtype=torch.int32))
    plans[prec] = pl
names = ["init"] + sum([[f"enc{l}.node_msg", f"enc{l}.node_upd", f"enc{l}.edge_upd"] for l in range(3)], []) + sum([[f"dec{l}.msg", f"dec{l}.node_upd"] for l in range(3)], [])
def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
e0 = rel(plans["f16"].buffer("hE0").float(), plans["fp32"].buffer("hE0"))
print(f"hE0 rel {e0:.3e}")
for stop in range(1, len(names) + 1):
    out = {}
    for prec, pl in plans.items():
        pl.forward_partial(x, t, stop)
        torch.cuda.synchronize()
        out[prec] = {k: pl.buffer(k).float().cpu() for k in ("S", "hV", "hE", "out6")}
    nm = names[stop - 1]
    key = "S" if ("msg" in nm) else ("hE" if "edge_upd" in nm else "hV")
    a, b = out["f16"][key], out["fp32"][key]
    print(f"{stop:2d} {nm:14s} {key:3s} rel {rel(a, b):.3e}  maxabs {float((a-b).abs().max()):.3e}  ref max {float(b.abs().max()):.3e} nan {int(torch.isnan(a).sum())}")
a, b = out["f16"]["out6"], out["fp32"]["out6"]
print("out6 rel", rel(a, b), "maxabs", float((a - b).abs().max()))
