"""Where a train_latent step spends its time (one micro-batch of 16 ragged proteins): phases with CUDA events + wall clock, and the
kernel breakdown from torch.profiler."""
import sys, time, random, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, train, weights
from codlad_b200.diffusion import create_diffusion
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
rnd = random.Random(5000)
lens = sorted([rnd.randint(60, 400) for _ in range(128)], reverse=True)[:64]
prots = [synthetic.make_protein(n, 1, seed=5100 + i) for i, n in enumerate(lens)]
batch = synthetic.collate_many(prots)
tr = train.DenoiserTrainer(weights.init_denoiser_state(0), gemm=mode)
diffusion = create_diffusion("")
x1 = torch.randn(64, max(lens), 3)
t = torch.randint(0, 1000, (64,))
def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
geom = train.Geometry(batch, 64, tr.device)
print("edges", geom.E, "nodes", geom.Nn)
print("geometry ms", timed(lambda: train.Geometry(batch, 64, tr.device)))
xt = x1.cuda(); tt = t.cuda().float()
print("forward ms", timed(lambda: tr.forward(xt, tt, geom)))
def fb():
    out = tr.forward(xt, tt, geom); tr.zero_grad(); tr.backward(torch.ones_like(out))
print("forward+backward ms", timed(fb))
print("optimizer step ms", timed(lambda: tr.step()))
print("whole train_step ms", timed(lambda: tr.train_step(diffusion, x1, t, batch, dropout_p=0.6)))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fb(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
