"""Index build: same flags and outputs as the reference's retriever/index_text.py (:30-109):
encodes the collection, writes `<index_dir>/<checkpoint stem>.index` in faiss' IndexIDMap{IndexFlatIP}
layout and `meta.pkl`.  Rows are streamed to their final file offset batch by batch instead of being
held in host memory twice.

Launched under torchrun (`torchrun --nproc-per-node G index_text.py ...`) the build is sharded: rank r encodes rows
shard_ranges(N, G)[r] of the collection on its own GPU and writes them at their final offsets of the ONE index file
(cldrd_index_writer_open_range); rank 0 creates the file, gathers the ids and writes the id array and meta.pkl.  The
file is byte-identical to a single-process build (the 2.5 h the reference reports for 8.8 M passages on one GPU,
README.md:20, are encoder time: it divides by G).

An interrupted build can be continued (`--continue_build`): every rank records, in `<index file>.progress.<rank>of<G>`,
how many of its rows are on stable storage (fdatasync first, then the record, every CLDRD_BUILD_SYNC_ROWS rows); a
continued build skips those rows, encodes the rest in the same batches and ends with the same bytes as an
uninterrupted one.  The reference's build holds everything in RAM until the end and starts over."""
import argparse
import ctypes as C
import errno
import glob
import json
import os
import pickle
import sys
import zlib
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
    parser.add_argument("--continue_build", action="store_true", default=False,
                        help="continue an interrupted build of the same index file instead of starting over")
    args = parser.parse_args(argv)
    if args.resume:
        assert args.index_dir[:-7] in args.resume     # same guard as the reference (:50)
    os.makedirs(args.index_dir, exist_ok=True)     # (several ranks may get here at once)
    return args


def _free_bytes(directory) -> int:
    vfs = os.statvfs(directory)
    return vfs.f_bavail * vfs.f_frsize


def _read_progress(path, expect):
    """(rows of this rank's range already on stable storage, NaNs counted in them) from a progress record that belongs
    to the same build (same collection size, dimension, row range and batch size); (0, 0) otherwise."""
    try:
        with open(path) as f:
            p = json.load(f)
    except (OSError, ValueError):
        return 0, 0
    if not isinstance(p, dict) or any(p.get(k) != v for k, v in expect.items()):
        return 0, 0
    done = p.get("done", 0)
    if not isinstance(done, int) or not 0 <= done <= expect["nrows"]:
        return 0, 0
    return done, int(p.get("n_nan", 0))


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
    stem = (Path(args.resume).stem.split(".")[0] if args.resume else args.index_name) + ".index"
    index_path = os.path.join(args.index_dir, stem)
    hidden = model.query_encoder.config.hidden_size
    # Continue or start over: rank 0 looks at the file, everybody follows its decision.
    fresh = True
    if rank == 0:
        if args.continue_build and os.path.exists(index_path):
            # an interrupted build left headers + some rows and no id array yet: joining it as a writer of zero rows
            # succeeds exactly when the file declares this collection size, dimension and layout
            probe = C.c_void_p()
            if lib().cldrd_index_writer_open_range(C.byref(probe), index_path.encode(), n, hidden, 1, 0, 0, 0, 0) == 0:
                lib().cldrd_index_writer_finish(probe, None)
                fresh = False
        if fresh:
            for stale in glob.glob(glob.escape(index_path) + ".progress.*"):
                os.remove(stale)
            # fail now, not after hours of encoding: the whole file (headers + rows + id array) must fit
            need = 82 + n * hidden * 4 + 8 + n * 8
            have = _free_bytes(args.index_dir) + (os.path.getsize(index_path) if os.path.exists(index_path) else 0)
            if have < need:
                fresh = None
    if world > 1:
        import torch.distributed as dist
        box = [fresh]
        dist.broadcast_object_list(box, src=0)
        fresh = box[0]
    if fresh is None:      # every rank raises (nobody is left waiting in a collective)
        raise OSError(errno.ENOSPC, f"not enough room in {args.index_dir!r} for an index of {n} x {hidden} rows "
                                    f"({82 + n * hidden * 4 + 8 + n * 8} bytes)")
    progress_path = f"{index_path}.progress.{rank}of{world}"
    # what a progress record must agree on to be continued: same collection size and dimension, same row range, same
    # batches, and the same ids in the same order in this rank's range (a collection file that was edited in between)
    ids_crc = zlib.crc32(np.fromiter((dataset.id_seq_pair[i][0] for i in rr), dtype=np.int64, count=len(rr)).tobytes())
    expect = {"n": n, "d": hidden, "row0": rr.start, "nrows": len(rr), "batch_size": args.batch_size, "ids_crc": ids_crc,
              "is_query": bool(args.is_query), "max_length": int(args.max_length)}
    done, n_nan = (0, 0) if fresh else _read_progress(progress_path, expect)
    if done:
        print(f"[rank {rank}] continuing the build at row {rr.start + done} ({done} of {len(rr)} rows already in the file)")
    todo = range(rr.start + done, rr.stop)
    part = Subset(dataset, todo) if (world > 1 or done) else dataset
    # the reference tokenises with 4 worker processes (retriever/index_text.py:84)
    workers = int(os.environ.get("CLDRD_LOADER_WORKERS", "4"))
    loader = DataLoader(part, batch_size=args.batch_size, shuffle=False, num_workers=workers, collate_fn=dataset.collate_fn)
    w = C.c_void_p()
    creator = fresh and rank == 0      # creates the file and its headers; the others open it once it exists
    if creator:
        check(lib().cldrd_index_writer_open_range(C.byref(w), index_path.encode(), n, hidden, 1, 0, todo.start, len(todo), 1))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if not creator:
        check(lib().cldrd_index_writer_open_range(C.byref(w), index_path.encode(), n, hidden, 1, 0, todo.start, len(todo), 0))
    sync_rows = int(os.environ.get("CLDRD_BUILD_SYNC_ROWS", "65536"))        # 0: no progress records
    fault_after = int(os.environ.get("CLDRD_FAULT_BUILD_AFTER_ROWS", "0"))   # fault injection (tests): die after that many rows
    flushed = done                       # rows of this rank's range handed to the file
    recorded = done                      # ... of which on stable storage and in the progress record

    def record_progress():
        nonlocal recorded
        if sync_rows <= 0 or flushed == recorded:
            return
        check(lib().cldrd_index_writer_sync(w))          # the record never runs ahead of the data
        tmp = progress_path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(dict(expect, done=flushed, n_nan=n_nan), f)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, progress_path)
        recorded = flushed

    text_ids = [dataset.id_seq_pair[i][0] for i in range(rr.start, rr.start + done)]      # ids follow the file order
    # Two page-locked slots: batch i's rows travel device -> host and are appended to the file while the encoder
    # already works on batch i+1 (the reference blocks on `.cpu().numpy()` after every batch, retrieval_utils.py:47).
    on_gpu = dev.type == "cuda"
    slots = [torch.empty((args.batch_size, hidden), dtype=torch.float32, pin_memory=on_gpu) for _ in range(2)]
    ready = [torch.cuda.Event() if on_gpu else None for _ in range(2)]
    pending = []      # (slot, rows) copied but not yet written

    def flush_one():
        nonlocal n_nan, flushed
        slot, m = pending.pop(0)
        if on_gpu:
            ready[slot].synchronize()
        rows = slots[slot][:m].numpy()
        n_nan += int(np.isnan(rows).sum())
        check(lib().cldrd_index_writer_append(w, ptr(rows), m))
        flushed += m
        if sync_rows > 0 and flushed - recorded >= sync_rows:
            record_progress()
        if fault_after and flushed - done >= fault_after:
            raise RuntimeError(f"CLDRD_FAULT_BUILD_AFTER_ROWS: injected failure after {flushed - done} rows")

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
        record_progress()      # this rank's range is complete: a continued build has nothing left to encode here
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
            lib().cldrd_index_writer_finish(w, None)      # (and can be completed with --continue_build)
    if rank == 0:
        meta = {"text_ids": text_ids_arr, "text_id_to_idx": {tid: idx for idx, tid in enumerate(text_ids)}}
        with open(os.path.join(args.index_dir, "meta.pkl"), "wb") as f:
            pickle.dump(meta, f)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()       # every rank's rows are in the file when anybody returns
        if own_group:
            dist.destroy_process_group()
    if os.path.exists(progress_path):      # the file is complete (ids and meta.pkl included): nothing to continue
        os.remove(progress_path)
    return index_path


if __name__ == "__main__":
    main(get_args())
