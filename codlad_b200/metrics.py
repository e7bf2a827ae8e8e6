"""Sample-quality metrics of the evaluation step that follows sampling: the reference's
`eval_sample_qualities` (utils/protein_module.py:335-364) and `valid_ratio_and_cut_off_result` (test.py:168-188) on the
GPU.  The O(Na^2) bond-graph comparison and the RMSD sums run in libcodlad_b200.so (`cb2_eval_bond_graphs`); the pair-list
losses of test.py:97-146 reduce in `cb2_pair_losses`; xyz / internal-coordinate losses (test.py:148-166) are torch elementwise ops.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _native as N

EPS = 1e-7      # test.py:27
# COVCUTOFFTABLE (utils/protein_module.py:128-): covalent radii in Angstrom used for the bond cut-off; the elements proteins contain
COV_RADIUS = {1: 0.23, 6: 0.68, 7: 0.68, 8: 0.68, 15: 0.75, 16: 1.02, 34: 1.22}
MAX_Z = max(COV_RADIUS)
_RADII_DEV = {}


def _radius_table(extra=None) -> torch.Tensor:
    table = dict(COV_RADIUS)
    table.update(extra or {})
    t = torch.zeros(max(table) + 1, dtype=torch.float32)
    for z, r in table.items():
        t[z] = r
    return t


def bond_graph_stats(xyz_ref: torch.Tensor, xyz_gen: torch.Tensor, atomic_nums: torch.Tensor, num_atoms, scale: float = 1.3, radii=None):
    """Raw per-structure statistics.  xyz_* [sum Na, 3], atomic_nums [sum Na], num_atoms [n] -> (counts [n, 6] int64, sums [n, 4] f64)
    on the device: counts = {differing adjacency entries, reference bonds, generated bonds} for all atoms, then heavy atoms only;
    sums = {sum of squared deviations, Na, the same over heavy atoms, number of heavy atoms}."""
    N.require_cuda()
    dev = xyz_gen.device if xyz_gen.is_cuda else torch.device("cuda")
    ref = xyz_ref.to(dev, torch.float32).contiguous()
    gen = xyz_gen.to(dev, torch.float32).contiguous()
    z = atomic_nums.to(dev, torch.int32).contiguous()
    num = torch.as_tensor(num_atoms, dtype=torch.int64).cpu()
    if ref.shape != gen.shape or ref.shape[0] != z.shape[0] or int(num.sum()) != ref.shape[0]:
        raise ValueError("xyz_ref / xyz_gen / atomic_nums / num_atoms do not describe the same atoms")
    unknown = set(int(v) for v in torch.unique(z).tolist()) - set(COV_RADIUS) - set(radii or {})
    if unknown:
        raise KeyError(f"no covalent radius for atomic numbers {sorted(unknown)}")
    offsets = torch.zeros(num.numel() + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(num, 0)
    key = (str(dev), tuple(sorted((radii or {}).items())))
    rad = _RADII_DEV.get(key)
    if rad is None:
        rad = _RADII_DEV[key] = _radius_table(radii).to(dev)
    n = int(num.numel())
    counts = torch.empty(n, 6, dtype=torch.int64, device=dev)
    sums = torch.empty(n, 4, dtype=torch.float64, device=dev)
    off_dev = offsets.to(dev, non_blocking=True)
    N.check(N.lib().cb2_eval_bond_graphs(N.dptr(ref), N.dptr(gen), N.dptr(z), N.dptr(off_dev), n, int(num.max()) if n else 0, N.dptr(rad),
                                         int(rad.numel() - 1), C.c_float(scale), N.dptr(counts), N.dptr(sums), N.stream_ptr()), "eval_bond_graphs")
    return counts, sums


def eval_sample_qualities(xyz_ref, xyz_gen, atomic_nums, num_atoms, scale: float = 1.3):
    """Per structure (one generated structure per reference structure, as test.py:178-181 calls it): dict of tensors
    `heavy_valid`, `all_valid` (bond graph identical to the reference's), `heavy_graph_diff_ratio`, `all_graph_diff_ratio`
    (|#ref bonds - #gen bonds| / #ref bonds, protein_module.py:315), `all_rmsd`, `heavy_rmsd` (protein_module.py:325-333)."""
    counts, sums = bond_graph_stats(xyz_ref, xyz_gen, atomic_nums, num_atoms, scale)
    c = counts.to(torch.float64)
    return {
        "all_valid": counts[:, 0] == 0, "heavy_valid": counts[:, 3] == 0,
        "all_graph_diff_ratio": (c[:, 1] - c[:, 2]).abs() / c[:, 1], "heavy_graph_diff_ratio": (c[:, 4] - c[:, 5]).abs() / c[:, 4],
        "all_rmsd": torch.sqrt(sums[:, 0] / sums[:, 1]), "heavy_rmsd": torch.sqrt(sums[:, 2] / sums[:, 3]),
    }


def valid_ratio_and_cut_off_result(xyz, xyz_recon, num_atoms, atomic_nums):
    """test.py:168-188: four per-structure lists (heavy valid ratio, all-atom valid ratio, heavy graph-difference ratios,
    all-atom graph-difference ratios); each structure is compared with its single reconstruction, so a ratio is 0.0 or 1.0 and
    the difference lists hold one-element lists."""
    q = eval_sample_qualities(xyz, xyz_recon, atomic_nums, num_atoms)
    hv, av = q["heavy_valid"].cpu().tolist(), q["all_valid"].cpu().tolist()
    hg, ag = q["heavy_graph_diff_ratio"].cpu().tolist(), q["all_graph_diff_ratio"].cpu().tolist()
    return [float(v) for v in hv], [float(v) for v in av], [[float(v)] for v in hg], [[float(v)] for v in ag]


def superposed_rmsd(xyz_a: torch.Tensor, xyz_b: torch.Tensor, num_atoms) -> torch.Tensor:
    """Minimum RMSD under rigid superposition of structure pairs (md.rmsd as test.py:37-79 uses it).  xyz_* [sum Na, 3] (or
    [n, Na, 3] with num_atoms = None) -> [n] float64 on the device (`cb2_superposed_rmsd`)."""
    N.require_cuda()
    dev = xyz_a.device if xyz_a.is_cuda else torch.device("cuda")
    if num_atoms is None:
        num_atoms = [xyz_a.shape[1]] * xyz_a.shape[0]
    a = xyz_a.to(dev, torch.float32).reshape(-1, 3).contiguous()
    b = xyz_b.to(dev, torch.float32).reshape(-1, 3).contiguous()
    num = torch.as_tensor(num_atoms, dtype=torch.int64).cpu()
    if a.shape != b.shape or int(num.sum()) != a.shape[0]:
        raise ValueError("superposed_rmsd: the two sets do not describe the same atoms")
    offsets = torch.zeros(num.numel() + 1, dtype=torch.int64)
    offsets[1:] = torch.cumsum(num, 0)
    out = torch.empty(num.numel(), dtype=torch.float64, device=dev)
    N.check(N.lib().cb2_superposed_rmsd(N.dptr(a), N.dptr(b), N.dptr(offsets.to(dev)), int(num.numel()), N.dptr(out), N.stream_ptr()), "superposed_rmsd")
    return out


def compute_div(gen_structures, ref_structure):
    """test.py:37-96 (compute_rmsd_ref, compute_rmsd_gen, compute_div): DIV = 1 - mean rmsd(gen, mean gen) / mean rmsd(gen, ref) over the
    G generated ensembles `gen_structures` (list of [P, Na, 3]) of P structures and their reference [P, Na, 3]; every RMSD is the
    superposed one.  Returns (div, rmsd_ref, rmsd_gen) as Python floats."""
    gen = torch.stack([torch.as_tensor(g, dtype=torch.float32) for g in gen_structures], 0)          # [G, P, Na, 3]
    ref = torch.as_tensor(ref_structure, dtype=torch.float32)
    G, P, Na, _ = gen.shape
    dev = gen.device if gen.is_cuda else torch.device("cuda")
    gen, ref = gen.to(dev), ref.to(dev)
    mean_gen = gen.mean(0)                                                                            # np.mean(gen_structures, axis=0)
    flat = gen.reshape(G * P, Na, 3)
    rmsd_ref = superposed_rmsd(flat, ref[None].expand(G, -1, -1, -1).reshape(G * P, Na, 3), None).mean()
    rmsd_gen = superposed_rmsd(flat, mean_gen[None].expand(G, -1, -1, -1).reshape(G * P, Na, 3), None).mean()
    return float(1.0 - rmsd_gen / rmsd_ref), float(rmsd_ref), float(rmsd_gen)


# -- pair-list evaluation losses (test.py:97-151) -------------------------------------------------------------------------------
def pair_losses(xyz, idx, xyz_data=None, thr_count: float = 0.0, thr_hinge: float = 0.0) -> torch.Tensor:
    """`cb2_pair_losses`: [4] float64 on the device = {#(d < thr_count), sum max(d - thr_hinge, 0), sum (d - d_data)^2, rows} over
    the pair ([P, 2]) or centre-pair ([P, 4]) list `idx`."""
    N.require_cuda()
    dev = xyz.device if xyz.is_cuda else torch.device("cuda")
    x = xyz.to(dev, torch.float32).contiguous()
    xd = None if xyz_data is None else xyz_data.to(dev, torch.float32).contiguous()
    ix = idx.to(dev, torch.int64).contiguous()
    if ix.dim() != 2 or ix.shape[1] not in (2, 4):
        raise ValueError("idx must be [P, 2] or [P, 4]")
    if ix.numel() and (int(ix.min()) < 0 or int(ix.max()) >= x.shape[0]):
        raise IndexError("pair index out of range")
    if xd is not None and xd.shape != x.shape:
        raise ValueError("xyz_data must have the shape of xyz")
    out = torch.empty(4, dtype=torch.float64, device=dev)
    N.check(N.lib().cb2_pair_losses(N.dptr(x), N.dptr(xd) if xd is not None else None, N.dptr(ix) if ix.numel() else None, int(ix.shape[1]),
                                    int(ix.shape[0]), C.c_float(thr_count), C.c_float(thr_hinge), N.dptr(out), N.stream_ptr()), "pair_losses")
    return out


def rows_once(a: torch.Tensor, b: torch.Tensor, n_atoms: int) -> torch.Tensor:
    """Rows of cat(a, b) ([*, 2] index pairs) that occur exactly once, in lexicographic order -- `uniques[counts == 1]` of
    test.py:120-123 -- on the device: int64 keys, one sort, `cb2_keys_once`."""
    N.require_cuda()
    dev = a.device if a.is_cuda else torch.device("cuda")
    rows = torch.cat((a.to(dev, torch.int64), b.to(dev, torch.int64)))
    keys = torch.sort(rows[:, 0] * int(n_atoms) + rows[:, 1]).values.contiguous()
    once = torch.empty(keys.numel(), dtype=torch.uint8, device=dev)
    N.check(N.lib().cb2_keys_once(N.dptr(keys) if keys.numel() else None, int(keys.numel()), N.dptr(once) if keys.numel() else None,
                                  N.stream_ptr()), "keys_once")
    k = keys[once.bool()]
    return torch.stack((k // int(n_atoms), k % int(n_atoms)), 1)


def ged_result(xyz_recon, xyz, edge_list):
    """test.py:141-146: mean squared difference of the bonded distances."""
    o = pair_losses(xyz_recon, edge_list, xyz_data=xyz)
    return (o[2] / o[3]).float()


def xyz_result(xyz_recon, xyz):
    """test.py:148-151."""
    return (xyz_recon - xyz).pow(2).sum(-1).mean()


def clash_result(edge_list, nbr_list, xyz_recon, bb_NO_list, non_bonded=None):
    """test.py:118-139: fraction of non-bonded neighbour pairs (rows occurring once in cat(edge_list, nbr_list)) closer than
    1.2 A plus the same fraction over the backbone N-O pairs.  `non_bonded`: that pair list if the caller kept it from an earlier
    call (it depends on the topology only)."""
    if non_bonded is None:
        non_bonded = rows_once(edge_list, nbr_list, xyz_recon.shape[0])
    a = pair_losses(xyz_recon, non_bonded, thr_count=1.2)
    b = pair_losses(xyz_recon, bb_NO_list.reshape(-1, 2), thr_count=1.2)
    frac = lambda o: (o[0] / o[3]).float() if float(o[3]) > 0 else torch.zeros((), device=o.device)
    return frac(a) + frac(b)


def inter_result(interaction_list, pi_pi_list, xyz_recon):
    """test.py:97-116: hinge losses on the side-chain interaction pairs (beyond 4 A) and on the pi-pi ring-centre pairs (beyond 6 A),
    weighted by their share of the interactions.  Returns (loss_inter, loss_pi_pi) like the reference (loss_inter includes the
    weighted pi-pi term)."""
    n_i, n_p = int(interaction_list.shape[0]), int(pi_pi_list.shape[0])
    dev = xyz_recon.device if xyz_recon.is_cuda else torch.device("cuda")
    loss_inter = torch.zeros((), device=dev)
    loss_pp = torch.zeros((), device=dev)
    if n_i > 0:
        o = pair_losses(xyz_recon, interaction_list, thr_hinge=4.0)
        loss_inter = (o[1] / o[3]).float() * (n_i / (n_i + n_p))
    if n_p > 0:
        o = pair_losses(xyz_recon, pi_pi_list, thr_hinge=6.0)
        loss_pp = (o[1] / o[3]).float()
        loss_inter = loss_inter + loss_pp * (n_p / (n_i + n_p))
    return loss_inter, loss_pp


def recon_result(ic_recon, ic, mask_):
    """test.py:153-166: bond MSE and chord-length angle / torsion errors over the real atoms."""
    m = torch.cat([mask_])
    n = m.sum()
    bond = ((ic_recon[:, :, 0] - ic[:, :, 0]).reshape(-1) * m).pow(2).sum() / n
    chord = lambda k: ((2 * (1 - torch.cos(ic[:, :, k] - ic_recon[:, :, k])) + EPS).sqrt().reshape(-1) * m).sum() / n
    return bond, chord(1), chord(2)
