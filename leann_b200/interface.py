"""Host-side mirror of LEANN's backend plugin API for the search path.

When the real ``leann`` core package is importable its classes are used as-is (so that
``@register_backend`` lands in LEANN's own ``BACKEND_REGISTRY`` and ``LeannSearcher`` finds
this backend by the ``backend_name`` stored in ``<index>.meta.json``).  On a box without
LEANN (the GPU test box) the same names are defined here with the same signatures:

  LeannBackendBuilderInterface / LeannBackendSearcherInterface / LeannBackendFactoryInterface
      packages/leann-core/src/leann/interface.py:7-20, 23-91, 94-107
  BACKEND_REGISTRY / register_backend
      packages/leann-core/src/leann/registry.py:16-27
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Literal, Optional

import numpy as np

try:  # pragma: no cover - exercised only where LEANN itself is installed
    from leann.interface import (  # type: ignore
        LeannBackendBuilderInterface,
        LeannBackendFactoryInterface,
        LeannBackendSearcherInterface,
    )
    from leann.registry import BACKEND_REGISTRY, register_backend  # type: ignore

    HAVE_LEANN = True
except Exception:  # ImportError or a broken partial install
    HAVE_LEANN = False

    class LeannBackendBuilderInterface(ABC):
        @abstractmethod
        def build(self, data: np.ndarray, ids: list[str], index_path: str, **kwargs) -> None: ...

    class LeannBackendSearcherInterface(ABC):
        @abstractmethod
        def __init__(self, index_path: str, **kwargs): ...

        @abstractmethod
        def _ensure_server_running(self, passages_source_file: str, port: Optional[int], **kwargs) -> int: ...

        @abstractmethod
        def search(self, query: np.ndarray, top_k: int, complexity: int = 64, beam_width: int = 1,
                   prune_ratio: float = 0.0, recompute_embeddings: bool = False,
                   pruning_strategy: Literal["global", "local", "proportional"] = "global",
                   zmq_port: Optional[int] = None, **kwargs) -> dict[str, Any]: ...

        @abstractmethod
        def compute_query_embedding(self, query: str, use_server_if_available: bool = True,
                                    zmq_port: Optional[int] = None) -> np.ndarray: ...

    class LeannBackendFactoryInterface(ABC):
        @staticmethod
        @abstractmethod
        def builder(**kwargs) -> LeannBackendBuilderInterface: ...

        @staticmethod
        @abstractmethod
        def searcher(index_path: str, **kwargs) -> LeannBackendSearcherInterface: ...

    BACKEND_REGISTRY: dict[str, Any] = {}

    def register_backend(name: str):
        def decorator(cls):
            BACKEND_REGISTRY[name] = cls
            return cls

        return decorator
