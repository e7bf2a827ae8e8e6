"""A tensor-core tier with fp32 STORAGE for the denoiser of the sampling loop: the strict (1e-3 A) tier of the path at TF32 arithmetic,
which is what the reference itself computes on a GPU (test.py:28-29, torch.backends.cuda.matmul.allow_tf32 = True).

It is the training forward of codlad_b200/train.py (the cb2t_* operators of include/codlad_b200_train.h: tcgen05.mma kind::tf32 GEMMs on
fp32 tensors, exact erf-GELU, fp32 LayerNorm / adaLN / neighbour sums) run without dropout, with everything that depends only on the
C-alpha trace (k-NN graph, raw edge features, their projections h_E0, the W_s term) hoisted out of the step loop, one diffusion step
(forward + the cb2_p_sample update) captured as a CUDA graph and replayed per step.  Unlike the f16 tier (edge_tc.cu / node_tc.cu) the
layers are NOT fused: every operator is one pass over h_E in HBM, so the tier is HBM-bound at about a fifth of the f16 tier's speed --
and 3-4x the SIMT fp32 tier it replaces as the accurate option.
"""
from __future__ import annotations

import torch

from . import _native as N
from .train import DenoiserTrainer, Geometry, _p


class Tf32Denoiser:
    def __init__(self, state_dict: dict, k_neighbors: int = 64):
        self.tr = DenoiserTrainer(state_dict, k_neighbors=k_neighbors, gemm="tf32")
        self.device = self.tr.device
        self.k_neighbors = int(k_neighbors)
        self.geom = None
        self._graph = None

    # -- per frame set -------------------------------------------------------------------------------------------------------
    def set_frames(self, batch: dict, n_members: int):
        """batch: reference-schema dict of F frames; the NB = n_members batch rows use frame b % F (ensemble-major, as test.py's doubled
        batch and sampler.frames_from_batch order them)."""
        F = int(batch["num_CGs"].numel())
        if n_members % F != 0:
            raise ValueError(f"{n_members} rows over {F} frames")
        reps = n_members // F
        batch = {"CG_nxyz": batch["CG_nxyz"].repeat(reps, 1), "num_CGs": batch["num_CGs"].repeat(reps)}
        self.geom = Geometry(batch, self.k_neighbors, self.device)
        ctx = {}
        self.tr._set_mode()
        self.hE0 = self.tr.edge_features(self.geom, ctx)
        hS = self.tr.params["W_s.weight"][self.geom.cg_z.reshape(-1).long()].contiguous()
        self.hS2 = self.tr.ops.ew(3, hS, scale=2.0)
        self.tr._set_mode(False)
        self._graph = None

    def forward(self, x, t):
        """x [NB, L, 3], t [NB] (original 0..999 scale) -> [NB, L, 6]."""
        return self.tr.forward(x, t, self.geom, hE0=self.hE0, hS2=self.hS2, keep=False)

    # -- sampling loop ---------------------------------------------------------------------------------------------------------
    def sample(self, diffusion, x, step_noise, use_graph: bool = True):
        """In-place reverse diffusion of x [NB, L, 3] with step_noise [T, NB, L, 3] (noise[s] is used at step s), T = diffusion.num_timesteps
        (gaussian_diffusion.py:451-547 + respace.py:117-129)."""
        g = self.geom
        T = diffusion.num_timesteps
        coef = torch.from_numpy(diffusion.coef_table()).to(self.device)
        tmap = torch.tensor([float(v) for v in diffusion.timestep_map], device=self.device)
        NB = g.B
        rows = g.Nn
        st = dict(x=x.to(self.device, torch.float32).contiguous().clone(), t=torch.zeros(NB, device=self.device),
                  step=torch.zeros(NB, dtype=torch.int32, device=self.device), noise=torch.zeros(NB, g.L, 3, device=self.device),
                  nxt=torch.zeros(NB, g.L, 3, device=self.device))
        lib = N.lib()

        def one_step():
            out = self.forward(st["x"], st["t"]).contiguous()
            N.check(lib.cb2_p_sample(_p(st["x"]), _p(out), _p(st["noise"]), _p(coef), _p(st["step"]), g.L, rows, 3, _p(st["nxt"]), N.stream_ptr()), "p_sample")
            st["x"].copy_(st["nxt"])

        graph = None
        if use_graph:
            st["t"].fill_(float(tmap[T - 1])); st["step"].fill_(T - 1)
            keep = st["x"].clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                one_step()                                  # warm-up outside capture: workspaces and TMA descriptors get their sizes
            torch.cuda.current_stream().wait_stream(s)
            st["x"].copy_(keep)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                one_step()
            st["x"].copy_(keep)
        for step in range(T - 1, -1, -1):
            st["t"].fill_(float(tmap[step]))
            st["step"].fill_(step)
            st["noise"].copy_(step_noise[step])
            graph.replay() if graph is not None else one_step()
        x.copy_(st["x"])
        return x
