"""Parameter inventories (names + shapes) of the two networks on the sampling path, and a
deterministic random initialiser for synthetic runs.

The names/shapes are the reference's `state_dict()` contract:
  * denoiser: `MPNN_models['mpnn_diffusion'](input_size=3, ...)`, models/latent_model.py:78-165,
    108 tensors / 2 449 974 parameters;
  * VQ-VAE decode side: `VAE(5, 36, ..., vqdim=3)` keys used by `latent_decode`
    (models/vae_model.py:686-706, 759-764, 830-839): `quantize._codebook.embed`, `map_out.*`,
    `equivaraintconv.*` (IC_Decoder :414-465 or IC_Decoder_angle :318-373).

`init_*_state` is NOT the reference's initialiser: the reference zero-initialises every adaLN
projection (latent_model.py:155-165), which makes a random-init network output a constant, so
synthetic parity runs use small random adaLN weights instead.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

H = 128
LATENT_STATS = {
    # datasets/miu_and_sigma/{PED_N6,PDB_K3,Atlas_K4}_x_{mean,std}.pt (3-vectors, fp32 bit patterns)
    ("N6", "PED"): ([1.068959355354309, -0.8994553089141846, 0.5618639588356018],
                    [5.1957831382751465, 4.400951385498047, 5.270322799682617]),
    ("K3", "PDB"): ([-1.5160585641860962, 0.6747006773948669, -0.5968422293663025],
                    [8.262883186340332, 5.664480686187744, 6.969945907592773]),
    ("K4", "Atlas"): ([-0.29618993401527405, 1.7351123094558716, -0.05292452499270439],
                      [5.226162910461426, 7.113760948181152, 6.114980697631836]),
}


def denoiser_shapes(input_size: int = 3, vocab: int = 30) -> "OrderedDict[str, tuple]":
    s = OrderedDict()

    def linear(name, out_f, in_f, bias=True):
        s[name + ".weight"] = (out_f, in_f)
        if bias:
            s[name + ".bias"] = (out_f,)

    linear("t_embedder.mlp.0", H, 256)
    linear("t_embedder.mlp.2", H, H)
    linear("x_in", H, input_size)
    linear("features.embeddings.linear", 16, 66)
    linear("features.edge_embedding", H, 16 + 16 * 9 + 7, bias=False)
    s["features.norm_edges.weight"] = (H,)
    s["features.norm_edges.bias"] = (H,)
    linear("W_e", H, H)
    s["W_s.weight"] = (vocab, H)
    for l in range(3):
        p = f"encoder_layers.{l}"
        linear(p + ".W1", H, 3 * H)
        linear(p + ".W2", H, H)
        linear(p + ".W3", H, H)
        linear(p + ".W11", H, 3 * H)
        linear(p + ".W12", H, H)
        linear(p + ".W13", H, H)
        linear(p + ".dense.W_in", 4 * H, H)
        linear(p + ".dense.W_out", H, 4 * H)
        linear(p + ".adaLN_modulation.1", 9 * H, H)
    for l in range(3):
        p = f"decoder_layers.{l}"
        linear(p + ".W1", H, 4 * H)
        linear(p + ".W2", H, H)
        linear(p + ".W3", H, H)
        linear(p + ".dense.W_in", 4 * H, H)
        linear(p + ".dense.W_out", H, 4 * H)
        linear(p + ".adaLN_modulation.1", 6 * H, H)
    linear("W_out.linear", 2 * input_size, H)
    linear("W_out.adaLN_modulation.1", 2 * H, H)
    return s


def ic_decoder_shapes(angle_variant: bool = False, prefix: str = "equivaraintconv.") -> "OrderedDict[str, tuple]":
    s = OrderedDict()
    D = 40

    def linear(name, out_f, in_f):
        s[prefix + name + ".weight"] = (out_f, in_f)
        s[prefix + name + ".bias"] = (out_f,)

    s[prefix + "res_embed.weight"] = (25, 4)
    for b in range(4):
        linear(f"message_blocks.{b}.inv_dense.0", D, D)
        linear(f"message_blocks.{b}.inv_dense.1", D, D)
        linear(f"message_blocks.{b}.dist_embed.block.1", D, 15)
    for b in range(4):
        linear(f"dense_blocks.{b}.1", D, D)
        linear(f"dense_blocks.{b}.3", D, D)
    s[prefix + "backbone_dist.weight"] = (25, 3)
    s[prefix + "sidechain_dist.weight"] = (25, 10)
    linear("backbone_angle.1", 3, D)
    linear("backbone_angle.3", 3, 3)
    if angle_variant:
        linear("sidechain_angle.1", 10, D)
        linear("sidechain_angle.3", 10, 10)
    else:
        s[prefix + "sidechain_angle.weight"] = (25, 10)
    linear("backbone_torsion.1", 3, D + 3)
    linear("backbone_torsion.3", 3, 3)
    T = D + 10 if angle_variant else D
    for b in range(4):
        linear(f"sidechain_torsion_blocks.{b}.1", T, T)
        linear(f"sidechain_torsion_blocks.{b}.3", T, T)
    linear("final_torsion.1", 10, T)
    linear("final_torsion.3", 10, 10)
    return s


def vae_decode_shapes(angle_variant: bool = False, codebook_size: int = 4096) -> "OrderedDict[str, tuple]":
    s = OrderedDict()
    s["quantize._codebook.embed"] = (1, codebook_size, 3)
    s["map_out.weight"] = (36, 3)
    s["map_out.bias"] = (36,)
    s.update(ic_decoder_shapes(angle_variant))
    return s


def _fill(shapes, seed, special):
    sd = OrderedDict()
    for n, (name, shape) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 100003 + n)
        if name in special:
            sd[name] = special[name](shape, g)
        elif len(shape) >= 2:
            bound = math.sqrt(6.0 / (shape[-1] + shape[-2]))
            sd[name] = (torch.rand(*shape, generator=g) * 2 - 1) * bound
        else:
            sd[name] = (torch.rand(*shape, generator=g) * 2 - 1) * 0.05
    return sd


def init_denoiser_state(seed: int = 0, adaln_std: float = 0.02, input_size: int = 3) -> "OrderedDict[str, torch.Tensor]":
    shapes = denoiser_shapes(input_size)
    special = {}
    for name in shapes:
        if "adaLN_modulation" in name:
            special[name] = lambda shape, g: torch.randn(*shape, generator=g) * adaln_std
    special["features.norm_edges.weight"] = lambda shape, g: 1.0 + 0.1 * (torch.rand(*shape, generator=g) - 0.5)
    special["W_s.weight"] = lambda shape, g: torch.randn(*shape, generator=g)
    return _fill(shapes, seed, special)


def init_vae_decode_state(seed: int = 0, angle_variant: bool = False, stats=("N6", "PED"), codebook_size: int = 4096,
                          gain: float = 0.5) -> "OrderedDict[str, torch.Tensor]":
    shapes = vae_decode_shapes(angle_variant, codebook_size)
    mean, std = (torch.tensor(v) for v in LATENT_STATS[stats])
    special = {"quantize._codebook.embed": lambda shape, g: mean + std * torch.randn(*shape, generator=g)}
    for name in shapes:
        if name.endswith(("res_embed.weight", "backbone_dist.weight", "sidechain_dist.weight", "sidechain_angle.weight")) \
                and len(shapes[name]) == 2 and shapes[name][0] == 25:
            special[name] = lambda shape, g: torch.randn(*shape, generator=g)
    sd = _fill(shapes, seed + 17, special)
    for name, t in sd.items():
        if name.startswith("equivaraintconv.") and name.endswith(".weight") and name not in special:
            sd[name] = t * gain      # keeps random-init decoder angles O(1) instead of O(1e5)
    return sd
