"""Residue topology tables for internal-coordinate reconstruction.

Domain data restated from the reference's templates (utils/utils_ic.py:6-83 `core_atoms`,
`atom_order_list`; utils/protein_module.py:72-93 `RES2IDX`), in a compact form:
for every residue-type id, the number of heavy atoms (slot order O, N, C, CA, side chain...)
and, for side-chain atom s (slot 4+s), the three already-placed slots it is built from.

`build_info` produces the `(permute, atom_idx, atom_orders)` tuple that
`traj_to_info` (utils/protein_module.py:455-494) derives from a PDB topology, for a
synthetic sequence whose atoms are already in template order (permute = identity).
"""
from __future__ import annotations

import torch

# residue-type id -> (name, heavy-atom count, side-chain build triples "abc" in base-14 digits)
# ids follow RES2IDX (N H A G R M S I E L Y D V W Q K P F C T O=TPO B=SEP).
_T = {
    0: ("ASN", 8, "123 234 345 456"),
    1: ("HIS", 10, "123 234 345 345 756 568"),
    2: ("ALA", 5, "123"),
    3: ("GLY", 4, ""),
    4: ("ARG", 11, "123 234 345 456 567 678 789"),
    5: ("MET", 8, "123 234 345 456"),
    6: ("SER", 6, "123 234"),
    7: ("ILE", 8, "123 234 345 346"),
    8: ("GLU", 9, "123 234 345 456 567"),
    9: ("LEU", 8, "123 234 345 456"),
    10: ("TYR", 12, "123 234 345 345 657 578 789 789"),
    11: ("ASP", 8, "123 234 345 456"),
    12: ("VAL", 7, "123 234 345"),
    13: ("TRP", 14, "123 234 345 345 756 657 579 79a a97 97c"),
    14: ("GLN", 9, "123 234 345 456 567"),
    15: ("LYS", 9, "123 234 345 456 567"),
    16: ("PRO", 7, "123 134 431"),
    17: ("PHE", 11, "123 234 345 456 567 345 459"),
    18: ("CYS", 6, "123 234"),
    19: ("THR", 7, "123 234 345"),
    20: ("TPO", 11, "123 234 234 645 457 457 457"),
    21: ("SEP", 10, "123 234 345 456 456 456"),
}

NUM_RESTYPES = len(_T)
SLOTS_PER_RESIDUE = 14
RES_NAMES = [_T[i][0] for i in range(NUM_RESTYPES)]
ATOM_COUNT = torch.tensor([_T[i][1] for i in range(NUM_RESTYPES)], dtype=torch.int64)


def _orders_table():
    tab = torch.empty(NUM_RESTYPES, 10, 3, dtype=torch.int64)
    tab[:, :, 0], tab[:, :, 1], tab[:, :, 2] = 0, 1, 2     # filler for absent atoms (protein_module.py:481-484)
    for rid, (_, _, spec) in _T.items():
        for s, tri in enumerate(spec.split()):
            tab[rid, s] = torch.tensor([int(ch, 16) for ch in tri])
    return tab


ORDERS = _orders_table()  # [22, 10, 3]


def build_info(restype: torch.Tensor):
    """restype [L] int64 (ids of the L reconstructed residues, termini already trimmed).
    Returns (permute [Na], atom_idx [Na], atom_orders [10, L, 3]) as int64 CPU tensors."""
    restype = restype.to(torch.int64).cpu()
    counts = ATOM_COUNT[restype]
    atom_idx = torch.cat([torch.arange(int(n)) + SLOTS_PER_RESIDUE * r for r, n in enumerate(counts)])
    permute = torch.arange(int(counts.sum()))
    atom_orders = ORDERS[restype].permute(1, 0, 2).contiguous()
    return permute, atom_idx, atom_orders


def slot_to_atom_map(info, num_residues: int) -> torch.Tensor:
    """Inverse of the reference's compaction `reshape(-1,3)[atom_idx][permute]`
    (utils/utils_ic.py:267): for each of the 14*L slots, the output atom row it lands in, or -1.
    Valid because `atom_idx[permute]` never repeats a slot (each atom is built once)."""
    permute, atom_idx, _ = info
    src = atom_idx[permute]
    inv = torch.full((SLOTS_PER_RESIDUE * num_residues,), -1, dtype=torch.int32)
    if src.numel() != torch.unique(src).numel():
        raise ValueError("info maps one slot to several atoms; not a valid topology")
    inv[src] = torch.arange(src.numel(), dtype=torch.int32)
    return inv
