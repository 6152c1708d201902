"""A module named `faiss` that exposes exactly the symbols CL-DRD's retriever touches, served by
libcldrd.so.  Put `cl-drd_b200/compat` on PYTHONPATH and the reference's own
retriever/retrieval_utils.py, index_text.py and retrieve_top_passages.py run unmodified.

The `cldrd` package is loaded from its file location instead of through sys.path: putting
`cl-drd_b200` itself on the path would shadow the reference's `retriever` package with ours."""
import importlib.util as _ilu
import os as _os
import sys as _sys

if "cldrd" not in _sys.modules:
    _pkg = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))), "cldrd")
    _spec = _ilu.spec_from_file_location("cldrd", _os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules["cldrd"] = _mod
    _spec.loader.exec_module(_mod)

from cldrd.index import (  # noqa: E402,F401
    METRIC_INNER_PRODUCT, METRIC_L2, GpuClonerOptions, GpuIndexFlat, GpuIndexShards,
    GpuMultipleClonerOptions, GpuResourcesVector, IndexFlatIP, IndexIDMap, IndexIDMap2, IntVector,
    StandardGpuResources, index_cpu_to_gpu, index_cpu_to_gpu_multiple, index_factory,
    index_gpu_to_cpu, read_index, write_index,
)

__version__ = "cldrd-b200"
