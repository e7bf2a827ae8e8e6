import sys, time, torch
sys.path.insert(0, '.')
import bench
from codlad_b200 import sampler, weights
torch.set_grad_enabled(False)
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
ens = int(sys.argv[2]) if len(sys.argv) > 2 else None
wl = bench.WORKLOADS[name]
if ens: wl = dict(wl, ensemble=ens)
if len(sys.argv) > 3: wl = dict(wl, L=int(sys.argv[3]))
if len(sys.argv) > 4: wl = dict(wl, k=int(sys.argv[4]))
bench.WORKLOADS[name] = wl
prot, batch, fs = bench._workload(0, name, 1)
print("frames", fs.F, "members", fs.NB, "L", fs.L, "csr edges", fs.csr_col.numel(), flush=True)
angle = wl["vae"] in ("K3", "K4")
bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0, angle, (wl["vae"], sampler.VAE_DATA[wl["vae"]])),
                        wl["vae"], k_neighbors=wl["k"], num_sampling_steps=100, precision="f16")
t0 = time.time(); plan = bm.upload(fs); torch.cuda.synchronize(); print("upload ok", time.time() - t0, flush=True)
x = torch.randn(fs.NB, fs.L, 3, device="cuda")
t0 = time.time(); out = plan.decode(bm.vae, x, denorm=True, num_atoms_total=fs.total_atoms); torch.cuda.synchronize(); print("decode ok", time.time() - t0, flush=True)
print("xyz finite", bool(torch.isfinite(out[3]).all()), flush=True)
tt = torch.full((fs.NB,), 999.0, device="cuda")
t0 = time.time(); o6 = plan.forward(x, tt); torch.cuda.synchronize(); print("forward ok", time.time() - t0, bool(torch.isfinite(o6).all()), flush=True)
plan.set_schedule(bm.diffusion.timestep_map, bm.diffusion.coef_table())
noise = torch.randn(100, fs.NB, fs.L, 3, device="cuda")
for rep in range(int(sys.argv[5]) if len(sys.argv) > 5 else 1):
    t0 = time.time(); plan.sample(x, noise, False); torch.cuda.synchronize(); print("eager sample ok", time.time() - t0, flush=True)
t0 = time.time(); plan.sample(x, noise, True); torch.cuda.synchronize(); print("graph sample ok", time.time() - t0, flush=True)
t0 = time.time(); plan.sample(x, noise, True); torch.cuda.synchronize(); print("graph sample 2 ok", time.time() - t0, flush=True)
