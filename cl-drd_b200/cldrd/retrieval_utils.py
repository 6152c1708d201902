"""Host helpers with the reference's names and semantics (retriever/retrieval_utils.py)."""
from __future__ import annotations

from timeit import default_timer as timer
from typing import List

import numpy as np

from . import index as _index


def index_retrieve(index, query_embeddings, topk, batch=None):
    """retriever/retrieval_utils.py:131-153.  batch=None -> one search, ndarrays back;
    batch=b -> slices of b queries, lists of lists back.  Returns (scores, neighbours).
    Prints the same three lines as the reference."""
    print("Query Num", len(query_embeddings))
    start = timer()
    if batch is None:
        nn_scores, nearest_neighbors = index.search(query_embeddings, topk)
    else:
        query_offset_base = 0
        nearest_neighbors: List[List[int]] = []
        nn_scores: List[List[float]] = []
        while query_offset_base < len(query_embeddings):
            batch_query_embeddings = query_embeddings[query_offset_base:query_offset_base + batch]
            batch_nn_scores, batch_nn = index.search(batch_query_embeddings, topk)
            nearest_neighbors.extend(batch_nn.tolist())
            nn_scores.extend(batch_nn_scores.tolist())
            query_offset_base += len(batch_query_embeddings)
    elapsed_time = timer() - start
    elapsed_time_per_query = 1000 * elapsed_time / max(len(query_embeddings), 1)
    print(f"Elapsed Time: {elapsed_time:.1f}s, Elapsed Time per query: {elapsed_time_per_query:.1f}ms")
    return nn_scores, nearest_neighbors


def index_retrieve_arrays(index, query_embeddings, topk):
    """Same search, but keeps (D, I) as ndarrays for the native run-file writer: one pass over
    the index for all queries instead of the reference's 128-query round trips."""
    print("Query Num", len(query_embeddings))
    start = timer()
    if not isinstance(query_embeddings, np.ndarray) and getattr(query_embeddings, "is_cuda", False):
        D, I = index.search(query_embeddings, topk)          # device-resident embeddings: no host round trip
    else:
        D, I = index.search(np.ascontiguousarray(query_embeddings, dtype=np.float32), topk)
    elapsed_time = timer() - start
    print(f"Elapsed Time: {elapsed_time:.1f}s, Elapsed Time per query: "
          f"{1000 * elapsed_time / max(len(query_embeddings), 1):.1f}ms")
    return D, I


def convert_index_to_gpu(index, faiss_gpu_index, useFloat16=False):
    """retriever/retrieval_utils.py:155-184: int -> one GPU, list -> row-sharded over the listed
    GPUs (the branch that raises NameError upstream works here)."""
    if type(faiss_gpu_index) == list and len(faiss_gpu_index) == 1:
        faiss_gpu_index = faiss_gpu_index[0]
    if isinstance(faiss_gpu_index, int):
        res = _index.StandardGpuResources()
        res.setTempMemory(1024 * 1024 * 1024)
        co = _index.GpuClonerOptions()
        co.useFloat16 = useFloat16
        return _index.index_cpu_to_gpu(res, faiss_gpu_index, index, co)
    assert isinstance(faiss_gpu_index, list)
    vres = _index.GpuResourcesVector()
    vdev = _index.IntVector()
    co = _index.GpuMultipleClonerOptions()
    co.shard = True
    co.useFloat16 = useFloat16
    for i in faiss_gpu_index:
        vdev.push_back(i)
        vres.push_back(_index.StandardGpuResources())
    return _index.index_cpu_to_gpu_multiple(vres, vdev, index, co)


def construct_flatindex_from_embeddings(embeddings, ids):
    """retriever/retrieval_utils.py:116-129."""
    hidden_size = embeddings.shape[1]
    print('embedding shape: ' + str(embeddings.shape))
    index = _index.index_factory(hidden_size, "Flat", _index.METRIC_INNER_PRODUCT)
    if ids is not None:
        if isinstance(ids, list):
            ids = np.array(ids)
        ids = ids.astype(np.int64)
        print(ids.shape, ids.dtype)
        index = _index.IndexIDMap2(index)
        index.add_with_ids(embeddings, ids)
    else:
        index.add(embeddings)
    return index


def get_embeddings_from_scratch(model, dataloader, use_fp16, is_query, show_progress_bar=False, to_device=False):
    """retriever/retrieval_utils.py:30-58: encoder forward under autocast, CLS vectors gathered to
    a float32 [N, hidden] ndarray + ids in file order.  The encoder stays a PyTorch module.

    to_device=True (extension) skips the reference's per-batch `.cpu().numpy()` round trip
    (retriever/retrieval_utils.py:47): the [N, hidden] float32 embeddings stay in HBM as one torch tensor, which
    `index.search` / `index_retrieve_arrays` accept directly."""
    import torch
    embeddings, embeddings_ids = [], []
    model.eval()
    dev = next(model.parameters()).device
    for batch in dataloader:
        with torch.no_grad():
            with torch.autocast(device_type=dev.type, dtype=torch.float16, enabled=bool(use_fp16) and dev.type == "cuda"):
                seq = {k: v.to(dev, non_blocking=True) for k, v in batch["seq"].items()}
                reps = model.query_embs(seq) if is_query else model.passage_embs(seq)
            text_ids = batch["id"]
        embeddings.append(reps.float() if to_device else reps.float().cpu().numpy())
        assert isinstance(text_ids, list)
        embeddings_ids.extend(text_ids)
    if to_device:
        embeddings = torch.cat(embeddings).contiguous() if embeddings else torch.empty((0, 0), device=dev)
        assert len(embeddings_ids) == embeddings.shape[0]
        print(f"# nan in embeddings: {int(torch.isnan(embeddings).sum().item())}")
        return embeddings, embeddings_ids
    embeddings = np.concatenate(embeddings)
    assert len(embeddings_ids) == embeddings.shape[0]
    print(f"# nan in embeddings: {np.sum(np.isnan(embeddings))}")
    return embeddings, embeddings_ids
