"""Host-side format conversion between the reference's ragged batch dict (CG_collate, reference
utils/dataset_module.py:259-295: rows of all frames concatenated, neighbour lists offset per sample) and the padded,
frame-major layout the CUDA plans use.  Vectorised torch indexing, no per-frame Python loops: the reference pays a
`.tolist()` sync plus a loop per call in reshape_and_create_mask / restore_shape (models/gcn_nn.py:35-52)."""
from __future__ import annotations

import torch


def ragged_positions(num: torch.Tensor):
    """num [F] lengths -> (frame [sum], pos [sum]) of every ragged row."""
    num = num.to(torch.int64)
    F = num.numel()
    frame = torch.repeat_interleave(torch.arange(F, device=num.device), num)
    off = torch.cumsum(num, 0) - num
    pos = torch.arange(int(frame.numel()), device=num.device) - off[frame]
    return frame, pos


def pad_frames(cg_nxyz: torch.Tensor, num: torch.Tensor, L: int | None = None):
    """CG_nxyz [sum L, 4] (col 0 = residue-type id, cols 1-3 = Angstrom) + num_CGs [F] ->
    X [F, L, 3] fp32 zero padded, cg_z [F, L] int32, on the device of `cg_nxyz`."""
    num = num.to(cg_nxyz.device, torch.int64)
    F = num.numel()
    L = int(num.max()) if L is None else L
    frame, pos = ragged_positions(num)
    cg = cg_nxyz.detach().to(torch.float32)
    X = torch.zeros(F, L, 3, device=cg.device)
    z = torch.zeros(F, L, dtype=torch.int32, device=cg.device)
    X[frame, pos] = cg[:, 1:4]
    z[frame, pos] = cg[:, 0].to(torch.int32)
    return X, z


def batch_csr(nbr: torch.Tensor, num: torch.Tensor, L: int):
    """CG_nbr_list [E, 2] with batch-global (ragged) node ids -> directed CSR over the F*L PADDED rows with
    frame-local column ids sorted ascending: (row_ptr [F*L+1] int32, col [E'] int32), on the device of `nbr` (a list that
    already lives on the GPU is converted there).  An undirected list (only i<j or only i>j pairs) is symmetrised first,
    exactly as make_directed does (reference models/gcn_nn.py:54-64)."""
    dev = nbr.device
    num = num.to(dev, torch.int64)
    F = num.numel()
    nbr = nbr.detach().to(torch.int64)
    if nbr.numel() == 0:
        return torch.zeros(F * L + 1, dtype=torch.int32, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)
    a, b = nbr[:, 0], nbr[:, 1]
    if not (bool((a > b).any()) and bool((b > a).any())):
        nbr = torch.cat([nbr, nbr.flip(1)], dim=0)
    ends = torch.cumsum(num, 0)
    off = ends - num
    frame = torch.bucketize(nbr[:, 0].contiguous(), ends, right=True)
    src = nbr[:, 0] - off[frame]
    dst = nbr[:, 1] - off[frame]
    row = frame * L + src
    order = torch.argsort(row * L + dst, stable=True)
    counts = torch.bincount(row, minlength=F * L)
    row_ptr = torch.zeros(F * L + 1, dtype=torch.int64, device=dev)
    row_ptr[1:] = torch.cumsum(counts, 0)
    return row_ptr.to(torch.int32), dst[order].to(torch.int32)


def merge_batches(batches):
    """Concatenate reference-schema batch dicts (one or more frames each) the way CG_collate does
    (utils/dataset_module.py:259-295): rows concatenated, CG_nbr_list shifted by the running residue count."""
    out, nbrs, off = {}, [], 0
    for b in batches:
        nbrs.append(b["CG_nbr_list"] + off)
        off += int(b["num_CGs"].sum())
    for k in batches[0]:
        if k == "CG_nbr_list":
            out[k] = torch.cat(nbrs, 0)
        elif torch.is_tensor(batches[0][k]):
            out[k] = torch.cat([b[k] for b in batches], 0)
    return out
