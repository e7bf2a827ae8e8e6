"""Needs a trace build: bash tools/dev/build_variant.sh trace -DCB2_TRACE_CTRL [-DCB2_TRACE_EPI]; CB2_LIB=codlad_b200/_variants/lib_trace.so"""
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
pl.buffer("tc_trace"); torch.cuda.synchronize()
pl.run_edge_kernel(mode, 1); torch.cuda.synchronize()
tr = pl.buffer("tc_trace").cpu().tolist()
ev = []
names = {"epi": ["E1 begin", "E1 acc ready", "E1 math done", "E1 handed off", "E2 begin", "E2 acc ready", "E2 math done", "E2 handed off", "drain begin", "drain acc ready", "drain done"],
         "mma": ["saw E1 done", "MMA2 issued", "saw E2 done", "saw drained", "saw load landed", "MMA1 issued", "warp start", "weights resident"],
         "tma": ["saw slot free", "load issued"]}
for who, base in (("epi", 0), ("mma", 5120), ("tma", 5632)):
    n = tr[base]
    for v in tr[base + 1: base + 1 + n]:
        v &= (1 << 64) - 1
        t, code = v >> 8, v & 0xff
        ev.append((t, who, names[who][code >> 2], code & 3))
ev.sort()
t0 = ev[0][0]
only = sys.argv[2] if len(sys.argv) > 2 else None
for t, who, name, g in ev[:int(sys.argv[3]) if len(sys.argv) > 3 else 200]:
    if only and who not in only.split(","): continue
    print(f"{(t - t0) / 1000:8.2f}  {who:3s} slot {g}  {name}")
