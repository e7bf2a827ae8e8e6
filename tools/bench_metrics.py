#!/usr/bin/env python
"""Times the evaluation-step kernel (cb2_eval_bond_graphs) on the ensemble configs[1] produces and the
CPU oracle port of the reference's eval_sample_qualities beside it.  Prints one JSON line.

    python tools/bench_metrics.py [--members 10] [--residues 300]
"""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codlad_b200 import metrics, sampler, synthetic, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=10)
    ap.add_argument("--residues", type=int, default=300)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    torch.set_grad_enabled(False)
    prot = synthetic.make_protein(args.residues, 1, seed=1002)
    fs = sampler.frames_from_batch(synthetic.collate(prot), prot.info, args.members)
    bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0), "N6", num_sampling_steps=10, precision="f16")
    xyz = bm.sample(bm.upload(fs), fs, generator=torch.Generator(device="cuda").manual_seed(1))["xyz"]
    na = int(fs.num_atoms[0])
    num = [na] * args.members
    ref = xyz[:na].repeat(args.members, 1).contiguous()           # every member against member 0
    z = torch.tensor([7, 6, 6, 8, 6, 6, 16, 8, 1], dtype=torch.int64).repeat(xyz.shape[0] // 9 + 1)[:xyz.shape[0]].cuda()
    for _ in range(3):
        metrics.bond_graph_stats(ref, xyz, z, num)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.reps):
        counts, sums = metrics.bond_graph_stats(ref, xyz, z, num)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.reps
    from oracle import restate as R
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    c_ref, _ = R.sample_quality_stats(ref.cpu(), xyz.cpu(), z.cpu(), num)
    cpu_s = time.perf_counter() - t0
    pairs = args.members * na * na
    print(json.dumps({"metric": "bond-graph comparisons (structure pairs scored) per second", "structures": args.members, "atoms_per_structure": na,
                      "gpu_ms_per_call": ms, "structures_per_s": args.members / (ms * 1e-3), "atom_pairs_per_s": 2 * pairs / (ms * 1e-3),
                      "cpu_port_s_per_call": cpu_s, "cpu_cores": os.cpu_count(), "speedup_vs_cpu_port": cpu_s / (ms * 1e-3),
                      "counts_equal_oracle": bool(torch.equal(counts.cpu(), c_ref)),
                      "note": "gpu_ms_per_call is the public call (input checks, one small H2D copy of the offsets, two kernels)"}))


if __name__ == "__main__":
    main()
