"""Decode side of the VQ-VAE behind the reference's call surface (reference models/vae_model.py:686-839,
utils/vq_module.py:98-163, utils/dataset_module.py:230-256): `quantize(x, mask=)`, `latent_decode(latent, mask, batch)`
and `get_norm_feature(..., norm_in=False)`, with `state_dict()` keys equal to the reference's for everything the
decode path reads (`quantize._codebook.embed`, `map_out.*`, `equivaraintconv.*`).

The e3nn encoder half (`encode`, `get_latent_wovq`, `get_latent_cg`) is out of scope for this path
(e3nn is not installed here and the encoder runs once per frame, not per step);
those methods raise.  All arithmetic happens in libcodlad_b200.so.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _native as N
from . import batching, weights
from .engine import Plan, VaeEngine
from .latent_model import _register
from .sampler import VAE_DATA


def get_norm_feature(feature, feature_type, norm_channel=True, norm_single=False, norm_in=True, dataname="PED"):
    """utils/dataset_module.py:230-256: latent (de)normalisation with the shipped mean / std vectors
    (datasets/miu_and_sigma/{dataname}_{feature_type}_x_{mean,std}.pt).  Same argument meaning as the reference:
    `feature_type` is the VAE type ('N6' / 'K3' / 'K4'), `dataname` the data set ('PED' / 'PDB' / 'Atlas'), as
    test.py:548 calls it -- get_norm_feature(samples, args.vae_type, norm_channel=..., norm_single=..., norm_in=False,
    dataname=args.data_type).  'IDRome_test_7' is remapped to the data set the VAE type was trained on (:239-246).
    The `_single` statistics files (norm_single=True) are not shipped with the reference and are not tabulated here."""
    if norm_single:
        raise NotImplementedError("get_norm_feature: the '_single' statistics (norm_single=True) are not shipped with the reference")
    if dataname == "IDRome_test_7":
        if feature_type not in VAE_DATA:
            raise KeyError(f"get_norm_feature: no data set for feature_type {feature_type!r}")
        dataname = VAE_DATA[feature_type]
    key = (feature_type, dataname)
    if key not in weights.LATENT_STATS:
        raise FileNotFoundError(f"get_norm_feature: no statistics for {dataname}_{feature_type}_x (known: "
                                f"{sorted(f'{d}_{v}_x' for v, d in weights.LATENT_STATS)})")
    mean, std = (torch.tensor(v, device=feature.device, dtype=feature.dtype) for v in weights.LATENT_STATS[key])
    return (feature - mean) / std if norm_in else feature * std + mean


class _Quantizer(nn.Module):
    """`vae.quantize(x, mask=)`: VectorQuantize.forward in eval mode (models/vae_model.py:740,835) -> (z_q, indices, loss).
    Holds the codebook under the reference's name `_codebook.embed`; masked-out positions return the input and index -1."""

    def __init__(self, owner):
        super().__init__()
        object.__setattr__(self, "_owner", owner)      # not a submodule: avoids a reference cycle in state_dict()

    def forward(self, x, mask=None):
        N.require_cuda()
        B, L, C = x.shape
        dev = torch.device("cuda", torch.cuda.current_device())
        xs = x.to(dev, torch.float32).contiguous()
        lengths = (mask.sum(1) if mask is not None else torch.full((B,), L)).to(dev, torch.int32).contiguous()
        if mask is not None and not bool((mask == (torch.arange(L, device=mask.device)[None] < mask.sum(1)[:, None])).all()):
            raise ValueError("quantize: mask must be a prefix mask (padding at the end), as reshape_and_create_mask produces")
        frame_of = torch.arange(B, device=dev, dtype=torch.int32)
        idx = torch.empty(B, L, device=dev, dtype=torch.int32)
        zq = torch.empty(B, L, C, device=dev, dtype=torch.float32)
        N.check(N.lib().cb2_vq_lookup(self._owner.engine().handle, N.dptr(xs), B, L, N.dptr(lengths), N.dptr(frame_of), 0, N.dptr(idx),
                                      N.dptr(zq), N.stream_ptr()), "vq_lookup")
        return zq, idx.long(), torch.zeros((), device=dev)


class VAE(nn.Module):
    """Decode-side VQ-VAE (N6: IC_Decoder; K3 / K4: IC_Decoder_angle)."""

    def __init__(self, vae_type: str = "N6", init_seed: int = 0):
        super().__init__()
        self.vae_type = vae_type
        self.angle_variant = vae_type in ("K3", "K4")
        stats = (vae_type, VAE_DATA[vae_type])
        self.quantize = _Quantizer(self)
        for name, t in weights.init_vae_decode_state(init_seed, self.angle_variant, stats).items():
            _register(self, name, t)
        self._engine = None
        self._decode_plans = {}
        self.eval()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Strict for the decode-side key set like the reference's load after remove_key (utils/model_module.py:118-120):
        a checkpoint that lacks any tensor `latent_decode` reads fails loudly instead of decoding with random weights.
        Encoder / prior tensors of a full checkpoint are ignored (that half is not built here); a DDP 'module.' prefix is
        stripped like the denoiser loader does."""
        state_dict = {k.removeprefix("module."): v for k, v in state_dict.items()}
        own = self.state_dict()
        picked = {k: v for k, v in state_dict.items() if k in own}
        if "quantize._codebook.embed" in picked:
            picked["quantize._codebook.embed"] = picked["quantize._codebook.embed"].reshape(own["quantize._codebook.embed"].shape)
        missing = [k for k in own if k not in picked]
        if strict and missing:
            raise RuntimeError(f"VAE.load_state_dict: {len(missing)} decode-side tensor(s) missing from the checkpoint: {missing[:5]}"
                               f"{'...' if len(missing) > 5 else ''}")
        out = super().load_state_dict(picked, strict=False, **kw)
        self.refresh()
        return out

    def refresh(self):
        for plan in self._decode_plans.values():
            plan.close()
        self._decode_plans.clear()
        if self._engine is not None:
            self._engine.close()
        self._engine = None

    def engine(self) -> VaeEngine:
        if self._engine is None:
            mean, std = weights.LATENT_STATS[(self.vae_type, VAE_DATA[self.vae_type])]
            self._engine = VaeEngine(self.state_dict(), mean, std, self.angle_variant)
        return self._engine

    # -- reference surface ---------------------------------------------------------------------------
    def latent_decode(self, latent, mask, batch):
        """vae_model.py:830-839: quantise -> map_out -> IC decoder.  latent [B, L, 3] (already de-normalised),
        mask [B, L] bool, batch = reference batch dict (CG_nxyz, num_CGs, CG_nbr_list).  -> (None, ic_recon [sum L, 13, 3]).
        The decode-only plan (buffers) is kept per geometry (B, L); the frame data are re-uploaded on every call."""
        num = batch["num_CGs"].to(torch.int64)
        B, L = latent.shape[0], latent.shape[1]
        if num.numel() != B or int(num.max()) != L:
            raise ValueError("latent_decode: latent must be padded to the batch's max length, one row per frame")
        X, z = batching.pad_frames(batch["CG_nxyz"], num, L)
        csr_row, csr_col = batching.batch_csr(batch["CG_nbr_list"], num, L)
        plan = self._decode_plans.get((B, L))
        if plan is None:
            if len(self._decode_plans) >= 4:
                self._decode_plans.pop(next(iter(self._decode_plans))).close()
            plan = self._decode_plans[(B, L)] = Plan(None, B, B, L, "fp32")
            plan._static = (torch.zeros(B, L + 2, 3), torch.zeros(B, L, 10, 3, dtype=torch.int8),
                            torch.full((B, L * 14), -1, dtype=torch.int32), torch.zeros(B, dtype=torch.int64))      # unused by the IC decoder
        plan.set_frames(X, num.to(torch.int32), z, torch.arange(B, dtype=torch.int32))
        ca, orders, slots, offs = plan._static
        plan.set_topology(self.engine(), ca, csr_row, csr_col, orders, slots, offs)
        _, _, ic, _ = plan.decode(self.engine(), latent, denorm=False, num_atoms_total=None, want_ic=True)
        m = mask.to(ic.device) if mask is not None else torch.ones(B, L, dtype=torch.bool, device=ic.device)
        return None, ic[m]                                         # restore_shape: ragged [sum L, 13, 3]

    def encode(self, *a, **k):
        raise NotImplementedError("the e3nn encoder (SURVEY.md section 8, row f-3) is outside this path")

    get_latent_wovq = get_latent = get_latent_cg = encode


def get_vae_model(modeltype="N6", modelpath=None, device=None, modelnum=None):
    """utils/model_module.py:20-123: -> (model, params).  Loads `modelpath` (a reference checkpoint / state_dict) when given."""
    model = VAE(modeltype)
    if modelpath is not None:
        sd = torch.load(modelpath, map_location="cpu")
        model.load_state_dict(sd.get("model", sd) if isinstance(sd, dict) else sd)
    return model, {"vae_type": modeltype, "codebook_size": 4096, "vqdim": 3, "cg_cutoff": 21.0}
