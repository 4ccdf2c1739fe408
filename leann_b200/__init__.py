"""leann_b200 — B200-native (sm_100a) implementation of LEANN's selective-recompute
search paths (HNSW and DiskANN/Vamana) behind LEANN's backend plugin API.  See DESIGN.md / INTEGRATION.md."""
from .interface import BACKEND_REGISTRY, register_backend  # noqa: F401

__all__ = ["BACKEND_REGISTRY", "register_backend", "capi", "backend", "csr", "synth"]


def __getattr__(name):
    # lazy: importing the package must not need the CUDA library (pure-host tools import csr/synth)
    import importlib

    if name in ("capi", "backend", "diskann_backend", "diskann_format", "vamana_build", "csr", "synth", "graph_build",
                "parallel", "build", "tooling"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
