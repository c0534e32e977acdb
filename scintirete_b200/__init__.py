"""scintirete_b200 — B200-native (sm_100a) search hot path of Scintirete behind the reference's
own index/distance interfaces. The product is `libscn_gpu.so` (CUDA, C ABI in include/scn_gpu.h);
this package is the tested host-side mirror of the Go cgo shim in go/."""
from .types import (DistanceMetric, ErrorCode, GraphState, HNSWParams, ScintireteError, SearchParams, SearchResult,
                    Vector)
from .index import (Batcher, DeviceStore, PinnedBuffer, DistanceCalculator, GPUFlatIndex, GPUHNSWIndex, IndexFactory, batch_distance,
                    dot_product, new_distance_calculator, normalize_vector, vector_magnitude)

__all__ = ["DistanceMetric", "ErrorCode", "GraphState", "HNSWParams", "ScintireteError", "SearchParams",
           "SearchResult", "Vector", "Batcher", "DeviceStore", "PinnedBuffer", "DistanceCalculator", "GPUFlatIndex", "GPUHNSWIndex",
           "IndexFactory", "batch_distance", "dot_product", "new_distance_calculator", "normalize_vector",
           "vector_magnitude"]
