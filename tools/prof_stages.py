#!/usr/bin/env python
"""Runs every stage of the path once inside a cudaProfilerStart/Stop window (after warm-up) so that

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/stages_<workload>.csv python tools/prof_stages.py <workload> [precision]

captures exactly those launches, in the order of bench.py's `roofline_all`; the number of kernels each stage launched is written to
gpurun_out/stages_<workload>.json.  tools/ncu_traffic.py joins the two into profiles/traffic.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from codlad_b200 import sampler, weights  # noqa: E402

torch.set_grad_enabled(False)
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
precision = sys.argv[2] if len(sys.argv) > 2 else "f16"
wl = bench.WORKLOADS[name]
prot, batch, fs = bench._workload(0, name, 1)
angle = wl["vae"] in ("K3", "K4")
bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0, angle, (wl["vae"], sampler.VAE_DATA[wl["vae"]])),
                        wl["vae"], k_neighbors=wl["k"], num_sampling_steps=10, precision=precision)
plan = bm.upload(fs)
bm.sample(plan, fs, generator=torch.Generator(device="cuda").manual_seed(1))          # realistic state in every buffer
xyz = torch.zeros(fs.total_atoms, 3, device="cuda")
stages = [
    ("edge message, encoder (edge kernel mode 0)", lambda: plan.run_stage(0, 1)),
    ("edge update, encoder (edge kernel mode 1)", lambda: plan.run_stage(1, 1)),
    ("edge message, decoder (edge kernel mode 2)", lambda: plan.run_stage(2, 1)),
    ("node update, encoder layer", lambda: plan.run_stage(3, 1)),
    ("node update, decoder layer", lambda: plan.run_stage(3, 4)),
    ("node update + FinalLayer + p_sample", lambda: plan.run_stage(3, 5)),
    ("k-NN graph", lambda: plan.run_stage(4)),
    ("edge featuriser", lambda: plan.run_stage(5)),
    ("IC distance filters", lambda: plan.run_stage(9, 0, bm.vae)),
    ("de-normalise + VQ lookup + map_out", lambda: plan.run_stage(6, 0, bm.vae)),
    ("IC decoder (messages + heads)", lambda: plan.run_stage(7, 0, bm.vae)),
    ("ic_to_xyz", lambda: plan.run_stage(8, 0, bm.vae, xyz)),
]
for _ in range(2):
    for _, fn in stages:
        fn()
torch.cuda.synchronize()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
counts = []
for nm, fn in stages:
    flush.fill_(1)                      # cold L2, like the `us_per_launch` column of roofline_all (the fill itself is outside the window)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    l0 = plan.launches
    fn()
    n = plan.launches - l0
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    counts.append((nm, max(1, int(n))))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"stages_{name}_{precision}.json"), "w") as f:
    json.dump(counts, f)
print("stages:", counts)
