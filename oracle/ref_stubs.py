"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Makes the *unmodified* reference (read-only checkout at /root/reference) importable
in the build container, where e3nn / torch_scatter / mdtraj / ase /
vector_quantize_pytorch are absent.  Used only by
`oracle/make_goldens.py` (fixture generation) and by the optional
`tests/test_oracle_vs_reference.py`, which is skipped when the checkout is absent
(it does not exist on the GPU box).

The stubs are empty namespaces except for two *functional* pieces the hot path
really calls:
  * torch_scatter.scatter_add   (models/vae_model.py:485-488) -> index_add_
  * vector_quantize_pytorch.VectorQuantize (utils/vq_module.py:106-111) -> the CPU
    restatement in oracle/restate.py (third-party, un-vendored, pinned ==1.21.7 in
    requirements.txt:34; parity for it is therefore UNPINNED).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CODLAD_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def install():
    """Register stub modules and put the reference on sys.path. Idempotent."""
    import torch

    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if "e3nn" not in sys.modules:
        e3nn = types.ModuleType("e3nn")
        o3 = types.ModuleType("e3nn.o3")
        e3nn.o3 = o3
        sys.modules["e3nn"] = e3nn
        sys.modules["e3nn.o3"] = o3
    if "mdtraj" not in sys.modules:
        sys.modules["mdtraj"] = types.ModuleType("mdtraj")
    if "ase" not in sys.modules:
        ase = types.ModuleType("ase")
        ase.Atoms = None
        sys.modules["ase"] = ase
    if "torch_scatter" not in sys.modules:
        ts = types.ModuleType("torch_scatter")

        def scatter_add(src, index, dim=0, dim_size=None):
            assert dim == 0
            out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
            return out.index_add_(0, index, src)

        ts.scatter_add = scatter_add
        ts.scatter = None
        ts.scatter_mean = None
        sys.modules["torch_scatter"] = ts
    if "vector_quantize_pytorch" not in sys.modules:
        from oracle import restate

        vq = types.ModuleType("vector_quantize_pytorch")
        vq.VectorQuantize = restate.VectorQuantizeEval
        for name in ("ResidualVQ", "GroupedResidualVQ", "RandomProjectionQuantizer", "FSQ", "LFQ"):
            setattr(vq, name, None)
        sys.modules["vector_quantize_pytorch"] = vq
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
