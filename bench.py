#!/usr/bin/env python
"""bench.py -- backmapped residues/sec of the reverse latent-diffusion sampling path.

One "step" = one pass of the whole hot path over one batch of synthetic input: BASELINE.json configs[1],
a PED-like 300-residue protein, N6 decoder, 100 respaced DDPM steps, num_ensemble = 10 -> 3000 backmapped
residues per rank per step (k-NN graph + edge features once per frame, 100 x denoiser forward + p_sample,
de-normalise, VQ lookup, IC decoder, ic_to_xyz).  At N > 1 every rank samples its own protein (weak scaling,
no collective on the data path).

    python bench.py --gpus N --steps K --warmup W [--precision f16|fp32] [--impl reference]

prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over ranks, L2 flushed
between timed steps), `e2e` the same metric through the public host API (pinned host inputs -> host coordinates),
`roofline` the dominant kernel (per-edge message MLP) timed alone with CUDA events, `cpu_baseline` the CPU oracle
port timed on the host cores on a bounded sample.  `--impl reference` times that CPU port as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "backmapped residues/sec (100-step, 10-ensemble)"
UNIT = "residues/s"
L_RES, ENSEMBLE, T_STEPS = 300, 10, 100
WORKLOAD = "configs[1]: PED-like 300-residue IDP, N6, 100-step latent sampling, num_ensemble=10"
# The bench line is configs[1] (`--workload c2`, the configuration the metric is quoted on).  c3 / c4 are BASELINE.json's larger
# configurations at their per-GPU shard size, runnable with the same harness for scale checks (not bench lines).
WORKLOADS = {
    "c2": dict(desc=WORKLOAD, L=300, frames=1, ensemble=10, k=64, compact=0.0, vae="N6"),
    "c3": dict(desc="configs[2] per-GPU shard: 32 PDB-like proteins x 500 residues, K3 decoder, 100 steps, num_ensemble=1",
               L=500, frames=32, ensemble=1, k=64, compact=0.0, vae="K3"),
    "c4": dict(desc="configs[3]: one Atlas-like 2000-residue frame, K4 decoder, k_neighbors=48, 100 steps, 32 members over the GPUs",
               L=2000, frames=1, ensemble=32, k=48, compact=0.0, vae="K4"),
}

# canonical algorithmic work per edge of the three per-edge kernels: GEMM flops only
EDGE_FLOPS = {0: 2 * 2 * 128 * 128, 1: 2 * 3 * 128 * 128, 2: 2 * 2 * 128 * 128}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _workload(rank: int, name: str = "c2", world: int = 1):
    """Synthetic protein(s) of this rank + host FrameSet (SURVEY.md section 8d generators)."""
    from codlad_b200 import sampler, synthetic
    w = WORKLOADS[name]
    if w["frames"] == 1:
        prot = synthetic.make_protein(w["L"], 1, seed=1002 + (rank if name == "c2" else 0), compact=w["compact"])
        batch = synthetic.collate(prot)
        ens = w["ensemble"] if name == "c2" else max(1, w["ensemble"] // world)      # c4 shards its members over the ranks
        return prot, batch, sampler.frames_from_batch(batch, prot.info, ens)
    prots = [synthetic.make_protein(w["L"], 1, seed=3000 + rank * w["frames"] + i, compact=w["compact"]) for i in range(w["frames"])]
    batch = synthetic.collate_many(prots)
    return prots[0], batch, sampler.frames_from_batch(batch, [p.info for p in prots], w["ensemble"])


# ------------------------------------------------------------------------------------------------ CPU port timing
def cpu_port_sample(denoiser_steps: int = 2, threads: int | None = None):
    """Times the CPU oracle (oracle/restate.py, a restatement of the reference's PyTorch path) on the full
    configs[1] batch for `denoiser_steps` denoiser+p_sample steps + graph/features + one decode + one ic_to_xyz and
    extrapolates the homogeneous steps to 100 (SURVEY.md section 8d / BASELINE.md section 3).  The reference
    itself is Python with absent third-party deps and cannot travel to the GPU box, hence kind = "port"."""
    from codlad_b200 import synthetic, weights
    from oracle import restate as R
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.set_grad_enabled(False)
    prot, batch, _ = _workload(0)
    dsd, vsd = weights.init_denoiser_state(0), weights.init_vae_decode_state(0)
    X = prot.ca_full[:, 1:-1].expand(ENSEMBLE, -1, -1).contiguous()
    zz = prot.restype_full[1:-1][None].expand(ENSEMBLE, -1).contiguous()
    mask = torch.ones(ENSEMBLE, L_RES, dtype=torch.bool)
    x = synthetic.latent_noise((ENSEMBLE, L_RES, 3), 1)
    noise = synthetic.latent_noise((denoiser_steps, ENSEMBLE, L_RES, 3), 2)
    sch = R.respaced_schedule(T_STEPS)
    t0 = time.perf_counter()
    g = R.edge_embedding(dsd, X[:1], mask[:1].int(), 64)           # once per frame (hoisted, like the CUDA path)
    t_graph = time.perf_counter() - t0
    graph = tuple(v.expand(ENSEMBLE, *v.shape[1:]).contiguous() for v in (g[0], g[3]))
    R.denoiser_forward(dsd, x, torch.full((ENSEMBLE,), 999), X, zz, mask, 64, graph=graph)      # warm-up
    t0 = time.perf_counter()
    for s in range(denoiser_steps):
        step = T_STEPS - 1 - s
        tt = torch.full((ENSEMBLE,), int(sch["timestep_map"][step]))
        out = R.denoiser_forward(dsd, x, tt, X, zz, mask, 64, graph=graph)
        x = R.p_sample_update(x, out, step, noise[s], sch)
    t_step = (time.perf_counter() - t0) / denoiser_steps
    mean, std = (torch.tensor(v) for v in weights.LATENT_STATS[("N6", "PED")])
    rep = lambda t: torch.cat([t] * ENSEMBLE, 0)
    nbr = torch.cat([batch["CG_nbr_list"] + e * L_RES for e in range(ENSEMBLE)], 0)
    t0 = time.perf_counter()
    ic, _ = R.latent_decode(vsd, x * std + mean, mask, rep(batch["CG_nxyz"][:, 0].long()), rep(batch["CG_nxyz"][:, 1:]), nbr,
                            torch.full((ENSEMBLE,), L_RES), False)
    og = batch["OG_CG_nxyz"].reshape(1, L_RES + 2, 4).expand(ENSEMBLE, -1, -1)
    R.ic_to_xyz(og, ic.reshape(ENSEMBLE, L_RES, 13, 3), prot.info)
    t_dec = time.perf_counter() - t0
    total = t_graph + T_STEPS * t_step + t_dec
    return {
        "value": L_RES * ENSEMBLE / total, "unit": UNIT, "cores": threads, "kind": "port",
        "sample": (f"full configs[1] batch (10x300): graph+features once ({t_graph:.2f}s) + {denoiser_steps} denoiser+p_sample steps "
                   f"({t_step:.2f}s each, extrapolated x{T_STEPS}) + 1 VQ/decode/ic_to_xyz ({t_dec:.2f}s); undoubled batch, "
                   f"features hoisted (the reference as shipped doubles the batch and recomputes features per step)"),
        "seconds_per_pass_extrapolated": total,
    }


def run_reference(args):
    """`--impl reference`: the CPU port on all host threads; each step is the bounded sample above."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_port_sample(denoiser_steps=1)
        if i >= args.warmup:
            vals.append(r)
    v = sum(x["value"] for x in vals) / len(vals)
    ms = 1e3 * sum(x["seconds_per_pass_extrapolated"] for x in vals) / len(vals)
    cb = dict(vals[-1]); cb["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU port of the reference path (oracle/restate.py); 100-step pass extrapolated from a bounded sample"},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ evaluation-step kernel
def run_metrics(args):
    """`--workload metrics`: the evaluation-step kernel (cb2_eval_bond_graphs) on the ensemble configs[1]
    produces, with the CPU oracle port of the reference's eval_sample_qualities timed beside it.  Not the bench line."""
    from codlad_b200 import metrics, sampler, synthetic, weights
    torch.set_grad_enabled(False)
    prot = synthetic.make_protein(L_RES, 1, seed=1002)
    fs = sampler.frames_from_batch(synthetic.collate(prot), prot.info, ENSEMBLE)
    bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0), "N6", num_sampling_steps=10, precision=args.precision)
    xyz = bm.sample(bm.upload(fs), fs, generator=torch.Generator(device="cuda").manual_seed(1))["xyz"]
    na = int(fs.num_atoms[0])
    num = [na] * ENSEMBLE
    ref = xyz[:na].repeat(ENSEMBLE, 1).contiguous()              # every member against member 0
    z = torch.tensor([7, 6, 6, 8, 6, 6, 16, 8, 1], dtype=torch.int64).repeat(xyz.shape[0] // 9 + 1)[:xyz.shape[0]].cuda()
    for _ in range(max(args.warmup, 3)):
        metrics.bond_graph_stats(ref, xyz, z, num)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 20)
    a.record()
    for _ in range(reps):
        counts, _ = metrics.bond_graph_stats(ref, xyz, z, num)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    from oracle import restate as R                               # CPU baseline leg
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    c_ref, _ = R.sample_quality_stats(ref.cpu(), xyz.cpu(), z.cpu(), num)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"metric": "bond-graph comparisons (structures scored per second)", "value": ENSEMBLE / (ms * 1e-3), "unit": "structures/s",
                      "structures": ENSEMBLE, "atoms_per_structure": na, "gpu_ms_per_call": ms, "atom_pairs_per_s": 2 * ENSEMBLE * na * na / (ms * 1e-3),
                      "cpu_baseline": {"value": ENSEMBLE / cpu_s, "unit": "structures/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": "the same 10 structures, oracle.restate.sample_quality_stats"},
                      "counts_equal_oracle": bool(torch.equal(counts.cpu(), c_ref)),
                      "note": "gpu_ms_per_call is the public call (input checks, one small H2D copy of the offsets, two kernels)"}))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="codlad_b200", choices=["codlad_b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CB2_PRECISION", "f16"), choices=["f16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["metrics"],
                    help="c2 = the bench line; c3 / c4 = scale checks; metrics = the evaluation-step kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "metrics":
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: codlad_b200 has no CPU fallback")
        return run_metrics(args)
    args.warmup = max(args.warmup, 3)

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: codlad_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.set_grad_enabled(False)
    from codlad_b200 import sampler, weights

    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    prot, batch, fs = _workload(rank, args.workload, world)
    fs.pin()
    angle = wl["vae"] in ("K3", "K4")
    bm = sampler.Backmapper(weights.init_denoiser_state(0), weights.init_vae_decode_state(0, angle, (wl["vae"], sampler.VAE_DATA[wl["vae"]])),
                            wl["vae"], k_neighbors=wl["k"], num_sampling_steps=T_STEPS, precision=args.precision)
    plan = bm.upload(fs)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: inputs already in HBM, per-step CUDA events, L2 flushed between steps
    for _ in range(args.warmup):
        bm.sample(plan, fs, generator=gen)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    l0 = plan.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        bm.sample(plan, fs, generator=gen)
        b.record()
    barrier()
    launches = plan.launches - l0
    ms_dev = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) / args.steps)
    residues = fs.NB * fs.L * world
    value = residues / (ms_dev * 1e-3)

    # ---- end to end through the host API: pinned host inputs -> H2D -> whole path -> D2H coordinates
    h2d = fs.host_bytes()
    for _ in range(2):
        xyz = bm.backmap_host(fs, generator=gen)
    d2h = xyz.numel() * xyz.element_size()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        bm.backmap_host(fs, generator=gen)
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    barrier()
    clk = clocks.stop() if clocks else None

    # ---- dominant kernel alone: encoder edge-update MLP (3 chained 128x128 GEMMs per edge + LN/adaLN epilogue)
    roof = None
    if rank == 0:
        peaks = _peaks()
        mode, reps = 1, 20
        edges = fs.NB * fs.L * plan.K
        for _ in range(3):
            plan.run_edge_kernel(mode, 1)
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); plan.run_edge_kernel(mode, 1); b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        ms_k = tot / reps
        achieved = EDGE_FLOPS[mode] * edges / (ms_k * 1e-3) / 1e12
        # the same kernel back to back without the flush: inside the step loop its input was written by the previous kernel and is
        # L2-resident, so this is the in-situ figure (reported next to the cold one, which stays the roofline entry)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            plan.run_edge_kernel(mode, 1)
        b.record()
        torch.cuda.synchronize()
        ms_warm = a.elapsed_time(b) / reps
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(f"edge_mode{mode}_{args.precision}")
        roof = {"kernel": f"edge kernel mode {mode} (encoder edge update), {args.precision} tier", "bound": "tensor",
                "achieved": achieved, "peak": peaks["tflops_burst"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops_burst"],
                "traffic": traffic, "peak_source": peaks["source"] + ", burst (kernel timed alone)", "us_per_launch": ms_k * 1e3,
                "us_per_launch_l2_resident": ms_warm * 1e3, "achieved_l2_resident": EDGE_FLOPS[mode] * edges / (ms_warm * 1e-3) / 1e12,
                "flops_per_launch": EDGE_FLOPS[mode] * edges, "algorithmic_bytes_per_launch": edges * 128 * 2 * (2 if args.precision == "f16" else 4)}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and args.workload == "c2":
        cpu = cpu_port_sample(denoiser_steps=2)

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 (tcgen05, fp32 accumulate)" if args.precision == "f16" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "residues_per_step_per_gpu": fs.NB * fs.L, "diffusion_steps": T_STEPS, "k_neighbors": plan.K,
                       "weights": "random init (seed 0), adaLN layers re-randomised", "l2": "flushed between timed steps (512 MiB fill)",
                       "sharding": f"{fs.F} frame(s) x {fs.NB // fs.F} member(s) per GPU, no collective"},
            "e2e": {"value": residues / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
