"""Value types of the hot path, mirrored from the reference's pkg/types/types.go and
internal/utils/errors.go so host code and tests read like the reference's own."""
from __future__ import annotations

import enum
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import numpy as np


class DistanceMetric(enum.IntEnum):
    """types.DistanceMetric (pkg/types/types.go:12-19); values cross the C ABI unchanged."""
    UNSPECIFIED = 0
    L2 = 1
    COSINE = 2
    INNER_PRODUCT = 3

    def __str__(self) -> str:  # types.go:22-33
        return {1: "L2", 2: "Cosine", 3: "InnerProduct"}.get(int(self), "Unspecified")


class ErrorCode(enum.IntEnum):
    """utils.ErrorCode subset used by this path (internal/utils/errors.go:11-49)."""
    INTERNAL = 1000
    RESOURCE = 1003
    VECTOR_NOT_FOUND = 3004
    DIMENSION_MISMATCH = 3005
    INVALID_PARAMETERS = 3007
    INDEX_BUILD_FAILED = 5000
    SEARCH_FAILED = 5001
    INSERT_FAILED = 5002


class ScintireteError(Exception):
    """utils.ScintireteError{Code, Message} (internal/utils/errors.go:128-160)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = int(code)
        self.message = message


@dataclass
class Vector:  # types.go:64-68
    id: int
    elements: np.ndarray
    metadata: Optional[Dict[str, Any]] = None


@dataclass
class SearchResult:  # types.go:83-86
    vector: Vector
    distance: float


@dataclass
class SearchParams:  # types.go:89-92
    top_k: int
    ef_search: Optional[int] = None


@dataclass
class HNSWParams:  # types.go:95-112 (defaults 16 / 200 / 50 / 16)
    m: int = 16
    ef_construction: int = 200
    ef_search: int = 50
    max_layers: int = 16
    seed: int = 0


@dataclass
class GraphState:
    """core.HNSWGraphState (interfaces.go:137-151), flattened: node i has id node_ids[i],
    list_counts[i] = len(Connections); its lists follow in layer order with edge_counts[...]
    neighbour ids each, concatenated in `edges`."""
    node_ids: np.ndarray
    list_counts: np.ndarray
    edge_counts: np.ndarray
    edges: np.ndarray
    entry_point: int
    max_layer: int
    size: int
    deleted: Optional[np.ndarray] = None
    vectors: Optional[np.ndarray] = None
    m: int = 16
    extra: Dict[str, Any] = field(default_factory=dict)
