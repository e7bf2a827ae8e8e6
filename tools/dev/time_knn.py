"""Dev: time cb2_knn_topk at the BASELINE shapes (eager launches, warm)."""
import sys, torch
sys.path.insert(0, '.')
from codlad_b200 import synthetic, _native as N
from codlad_b200.engine import knn_topk
for L, F, K, compact in ((300, 1, 64, 0.0), (500, 32, 64, 0.002), (700, 32, 64, 0.002), (2000, 1, 48, 0.003)):
    X = torch.stack([synthetic.make_protein(L, 1, seed=7 + f, compact=compact).ca_full[0, 1:-1] for f in range(F)]).cuda().contiguous()
    lengths = torch.full((F,), L, dtype=torch.int32, device="cuda")
    for _ in range(3): knn_topk(X, lengths, K)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): knn_topk(X, lengths, K)
    b.record(); torch.cuda.synchronize()
    print(f"L={L} F={F} K={K}: {a.elapsed_time(b) * 1e3 / 20:.1f} us per call (incl. two torch.empty)")
