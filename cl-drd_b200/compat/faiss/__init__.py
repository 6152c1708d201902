"""A module named `faiss` that exposes exactly the symbols CL-DRD's retriever touches, served by
libcldrd.so.  Put `cl-drd_b200/compat` (and `cl-drd_b200`) on PYTHONPATH and the reference's own
retriever/retrieval_utils.py, index_text.py and retrieve_top_passages.py run unmodified."""
import os as _os
import sys as _sys

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _pkg_root not in _sys.path:
    _sys.path.insert(0, _pkg_root)

from cldrd.index import (  # noqa: E402,F401
    METRIC_INNER_PRODUCT, METRIC_L2, GpuClonerOptions, GpuIndexFlat, GpuIndexShards,
    GpuMultipleClonerOptions, GpuResourcesVector, IndexFlatIP, IndexIDMap, IndexIDMap2, IntVector,
    StandardGpuResources, index_cpu_to_gpu, index_cpu_to_gpu_multiple, index_factory,
    index_gpu_to_cpu, read_index, write_index,
)

__version__ = "cldrd-b200"
