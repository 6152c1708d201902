"""Index build: same flags and outputs as the reference's retriever/index_text.py (:30-109):
encodes the collection, writes `<index_dir>/<checkpoint stem>.index` in faiss' IndexIDMap{IndexFlatIP}
layout and `meta.pkl`.  Rows are streamed to their final file offset batch by batch instead of being
held in host memory twice.

Launched under torchrun (`torchrun --nproc-per-node G index_text.py ...`) the build is sharded: rank r encodes rows
shard_ranges(N, G)[r] of the collection on its own GPU and writes them at their final offsets of the ONE index file
(cldrd_index_writer_open_range); rank 0 creates the file, gathers the ids and writes the id array and meta.pkl.  The
file is byte-identical to a single-process build (the 2.5 h the reference reports for 8.8 M passages on one GPU,
README.md:20, are encoder time: it divides by G)."""
import argparse
import ctypes as C
import os
import pickle
import sys
from pathlib import Path

import numpy as np
import torch
from torch.utils.data import DataLoader, Subset

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cldrd._lib import check, lib, ptr  # noqa: E402
from cldrd.encoder import DualEncoder, SequenceDataset, load_checkpoint  # noqa: E402
from cldrd.index import shard_ranges  # noqa: E402


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--resume", default="")
    parser.add_argument("--model_name_or_path", default="distilbert-base-uncased")
    parser.add_argument("--tokenizer_name_or_path", default="distilbert-base-uncased")
    parser.add_argument("--passages_path", default="collection.tsv")
    parser.add_argument("--queries_path", default="")
    parser.add_argument("--max_length", default=256, type=int)
    parser.add_argument("--index_dir", default="index/")
    parser.add_argument("--is_query", default=False, action="store_true")
    parser.add_argument("--is_parallel", default=True, type=lambda s: str(s).lower() not in ("0", "false", "no"))
    parser.add_argument("--share_weights", action="store_true", default=False)
    parser.add_argument("--batch_size", default=512, type=int)
    parser.add_argument("--index_name", default="", help="index file stem when --resume is empty")
    args = parser.parse_args(argv)
    if args.resume:
        assert args.index_dir[:-7] in args.resume     # same guard as the reference (:50)
    os.makedirs(args.index_dir, exist_ok=True)     # (several ranks may get here at once)
    return args


def main(args):
    from transformers import AutoTokenizer
    model = DualEncoder(args.model_name_or_path, share_weights=args.share_weights)
    print("************************* share weights = {} *************************".format(args.share_weights))
    if args.resume:
        print(f"load model from ==> {args.resume}")
        load_checkpoint(model, args.resume, args.is_parallel)
    elif not args.index_name:
        raise ValueError("not index path defined.")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    own_group = False
    if world > 1:
        import torch.distributed as dist
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        own_group = not dist.is_initialized()
        if own_group:
            if torch.cuda.is_available():
                dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
            else:
                dist.init_process_group("gloo")
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    model.to(dev).eval()
    tokenizer = AutoTokenizer.from_pretrained(args.tokenizer_name_or_path)
    path = args.queries_path if args.is_query else args.passages_path
    dataset = SequenceDataset.create_from_seqs_file(path, tokenizer, args.max_length, is_query=args.is_query)
    n = len(dataset)
    rr = shard_ranges(n, world)[rank]               # this rank's rows of the collection, in file order
    part = Subset(dataset, rr) if world > 1 else dataset
    # the reference tokenises with 4 worker processes (retriever/index_text.py:84)
    workers = int(os.environ.get("CLDRD_LOADER_WORKERS", "4"))
    loader = DataLoader(part, batch_size=args.batch_size, shuffle=False, num_workers=workers, collate_fn=dataset.collate_fn)
    stem = (Path(args.resume).stem.split(".")[0] if args.resume else args.index_name) + ".index"
    index_path = os.path.join(args.index_dir, stem)
    hidden = model.query_encoder.config.hidden_size
    w = C.c_void_p()
    if world > 1:
        import torch.distributed as dist
        if rank == 0:      # creates the file and its headers; the others open it once it exists
            check(lib().cldrd_index_writer_open_range(C.byref(w), index_path.encode(), n, hidden, 1, 0, rr.start, len(rr), 1))
        dist.barrier()
        if rank != 0:
            check(lib().cldrd_index_writer_open_range(C.byref(w), index_path.encode(), n, hidden, 1, 0, rr.start, len(rr), 0))
    else:
        check(lib().cldrd_index_writer_begin(C.byref(w), index_path.encode(), n, hidden, 1, 0))
    text_ids = []
    n_nan = 0
    # Two page-locked slots: batch i's rows travel device -> host and are appended to the file while the encoder
    # already works on batch i+1 (the reference blocks on `.cpu().numpy()` after every batch, retrieval_utils.py:47).
    on_gpu = dev.type == "cuda"
    slots = [torch.empty((args.batch_size, hidden), dtype=torch.float32, pin_memory=on_gpu) for _ in range(2)]
    ready = [torch.cuda.Event() if on_gpu else None for _ in range(2)]
    pending = []      # (slot, rows) copied but not yet written

    def flush_one():
        nonlocal n_nan
        slot, m = pending.pop(0)
        if on_gpu:
            ready[slot].synchronize()
        rows = slots[slot][:m].numpy()
        n_nan += int(np.isnan(rows).sum())
        check(lib().cldrd_index_writer_append(w, ptr(rows), m))

    finished = False
    try:
        for i, batch in enumerate(loader):
            with torch.no_grad():
                with torch.autocast(device_type=dev.type, dtype=torch.float16, enabled=on_gpu):
                    seq = {k: v.to(dev, non_blocking=True) for k, v in batch["seq"].items()}
                    reps = model.query_embs(seq) if args.is_query else model.passage_embs(seq)
            slot = i % 2
            while len(pending) >= 2 or any(p[0] == slot for p in pending):
                flush_one()
            m = reps.shape[0]
            slots[slot][:m].copy_(reps.float(), non_blocking=True)
            if on_gpu:
                ready[slot].record()
            pending.append((slot, m))
            text_ids.extend(batch["id"])
            if len(pending) == 2:
                flush_one()
        while pending:
            flush_one()
        if world > 1:      # ids of all ranks, in row order, on every rank (rank 0 writes them)
            import torch.distributed as dist
            parts = [None] * world
            dist.all_gather_object(parts, (text_ids, n_nan))
            text_ids = [t for p in parts for t in p[0]]
            n_nan = sum(p[1] for p in parts)
        if rank == 0:
            print(f"# nan in embeddings: {n_nan}")
            print("embs dtype: ", np.dtype(np.float32))
        text_ids_arr = np.array(text_ids, dtype=np.int64)
        check(lib().cldrd_index_writer_finish(w, ptr(text_ids_arr) if rank == 0 else None))
        finished = True
    finally:
        if not finished:      # encoder or I/O error mid-way: release the writer (handle, fd); the file stays incomplete
            lib().cldrd_index_writer_finish(w, None)
    if rank == 0:
        meta = {"text_ids": text_ids_arr, "text_id_to_idx": {tid: idx for idx, tid in enumerate(text_ids)}}
        with open(os.path.join(args.index_dir, "meta.pkl"), "wb") as f:
            pickle.dump(meta, f)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()       # every rank's rows are in the file when anybody returns
        if own_group:
            dist.destroy_process_group()
    return index_path


if __name__ == "__main__":
    main(get_args())
