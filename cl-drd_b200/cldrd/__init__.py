"""cldrd — B200-native exhaustive inner-product top-k search for CL-DRD's retrieval path.

Host side of libcldrd.so.  `cldrd.index` mirrors the faiss API surface the reference uses,
`cldrd.retrieval_utils` mirrors retriever/retrieval_utils.py, `cldrd.runfile` the run-file
writer, `cldrd.dist` the one-process-per-GPU sharded search (peer-memory exchange over NVLink, NCCL as fallback).
"""
from ._lib import CldrdError, LIB_PATH, lib  # noqa: F401
from .index import (  # noqa: F401
    METRIC_INNER_PRODUCT, METRIC_L2, GpuClonerOptions, GpuIndexFlat, GpuIndexShards,
    GpuMultipleClonerOptions, GpuResourcesVector, IndexFlatIP, IndexIDMap, IndexIDMap2, IntVector,
    StandardGpuResources, index_cpu_to_gpu, index_cpu_to_gpu_multiple, index_factory, read_index,
    shard_ranges, write_index,
)
from .runfile import RunFileStream, format_score, write_run_file  # noqa: F401

__version__ = "0.1.0"
from . import curriculum  # noqa: F401,E402
