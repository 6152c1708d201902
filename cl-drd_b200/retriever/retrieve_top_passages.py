"""Query search entry point: same flags, prints and run file as the reference's
retriever/retrieve_top_passages.py (:28-109).  The search runs on the B200 kernels; the regroup and
writer loops are one native call.

The embeddings stay on the device between the encoder and the search (no per-batch `.cpu().numpy()`
as in retriever/retrieval_utils.py:47).

Launched under torchrun (`torchrun --nproc-per-node G retrieve_top_passages.py ...`) it is the one-process-per-GPU
form: every rank loads only its row shard of the index file; rank 0 encodes the queries and broadcasts the
embeddings (the sharded protocol needs bit-identical replicas); all ranks take part in the sharded search
(cldrd.dist.ShardedSearcher.search); rank 0 writes the run file."""
import argparse
import os
import sys

import numpy as np
import torch
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cldrd  # noqa: E402
from cldrd.encoder import DualEncoder, SequenceDataset, load_checkpoint  # noqa: E402
from cldrd.retrieval_utils import convert_index_to_gpu, get_embeddings_from_scratch, index_retrieve_arrays  # noqa: E402


SEARCH_CHUNK = 8 * 8192     # queries per search call when the run file is streamed


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--resume", default="")
    parser.add_argument("--model_name_or_path", default="distilbert-base-uncased")
    parser.add_argument("--tokenizer_name_or_path", default="distilbert-base-uncased")
    parser.add_argument("--queries_path", default="queries.dev.small.tsv")
    parser.add_argument("--index_path", default="")
    parser.add_argument("--max_length", default=30, type=int)
    parser.add_argument("--top_k", default=1000, type=int)
    parser.add_argument("--is_parallel", default=True, type=lambda s: str(s).lower() not in ("0", "false", "no"))
    parser.add_argument("--share_weights", action="store_true", default=False)
    parser.add_argument("--output_path", default="")
    parser.add_argument("--gpus", default="0", help="comma-separated device list; more than one = row-sharded index")
    parser.add_argument("--precision", default="auto", choices=["auto", "f16", "bf16", "tf32", "simt"],
                        help="scan mode; results are exact fp32 in every mode")
    return parser.parse_args(argv)


def check_paths(queries_path, output_path):
    # retriever/retrieve_top_passages.py:48-59
    for needle, tag, msg in (("train", "train", "retrieve train queries"), ("dev", "dev", "retrieve dev queries"),
                             ("2019", "trec19", "retrieve trec-19 queries"), ("2020", "trec20", "retrieve trec-20 queries")):
        if needle in queries_path:
            print(msg)
            assert tag in output_path


def _loader(dataset):
    # the reference tokenises with 4 worker processes (retriever/retrieve_top_passages.py:81)
    workers = int(os.environ.get("CLDRD_LOADER_WORKERS", "4"))
    return DataLoader(dataset, batch_size=512, shuffle=False, num_workers=workers, collate_fn=dataset.collate_fn)


def _encode(args, is_query_side):
    from transformers import AutoTokenizer
    model = DualEncoder(args.model_name_or_path, share_weights=args.share_weights)
    print("************************* share weights = {} *************************".format(args.share_weights))
    if args.resume:
        print(f"load model from ==> {args.resume}")
        load_checkpoint(model, args.resume, args.is_parallel)
    model.cuda()
    tokenizer = AutoTokenizer.from_pretrained(args.tokenizer_name_or_path)
    dataset = SequenceDataset.create_from_seqs_file(args.queries_path, tokenizer, args.max_length, is_query=is_query_side)
    return get_embeddings_from_scratch(model, _loader(dataset), use_fp16=True, is_query=is_query_side,
                                       show_progress_bar=True, to_device=True)


def main(args, is_query_side=True, header="# unique query", guards=True):
    if guards:
        check_paths(args.queries_path, args.output_path)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    os.environ.setdefault("CLDRD_SCAN", args.precision)
    if world > 1:
        import torch.distributed as dist
        from cldrd.dist import ShardedSearcher
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local_rank)
        dev = torch.device("cuda", local_rank)
        own_group = not dist.is_initialized()
        if own_group:
            dist.init_process_group("nccl", device_id=dev)
        searcher = ShardedSearcher.from_file(args.index_path, device=local_rank, scan=args.precision)
        try:
            # one encoder run; every shard must see the same bits
            query_embs, query_ids = _encode(args, is_query_side) if rank == 0 else (None, None)
            shape = torch.tensor(list(query_embs.shape) if rank == 0 else [0, 0], dtype=torch.int64, device=dev)
            dist.broadcast(shape, src=0)
            if rank != 0:
                query_embs = torch.empty(tuple(shape.tolist()), dtype=torch.float32, device=dev)
            dist.broadcast(query_embs, src=0)
            # the run file grows batch by batch while later batches are still being searched
            stream = cldrd.RunFileStream(args.output_path) if rank == 0 else None
            qids = np.asarray(query_ids, dtype=np.int64) if rank == 0 else None
            searcher.search_to_host(query_embs, args.top_k,
                                    on_batch=(lambda b0, nb, D, I: stream.put(qids[b0:b0 + nb], I, D)) if rank == 0 else None)
            if rank == 0:
                print(f"{header} = {len(set(query_ids))}")
                avg = stream.close()
                print(f"average ranks per query = {avg}")
            dist.barrier()
        finally:
            searcher.close()
            if own_group:
                dist.destroy_process_group()
        return
    query_embs, query_ids = _encode(args, is_query_side)
    index = cldrd.read_index(args.index_path)                      # headers only; rows stream file -> HBM
    devs = [int(x) for x in str(args.gpus).split(",")]
    index = convert_index_to_gpu(index, devs if len(devs) > 1 else devs[0], False)
    if len(query_ids) <= SEARCH_CHUNK:
        nn_scores, nn_doc_ids = index_retrieve_arrays(index, query_embs, args.top_k)
        print(f"{header} = {len(set(query_ids))}")
        avg = cldrd.write_run_file(args.output_path, query_ids, nn_doc_ids, nn_scores)
    else:
        # large query sets (the 502 939 training queries of the curriculum step, :48-50): search chunk i+1 while the
        # writer thread formats chunk i
        stream = cldrd.RunFileStream(args.output_path)
        qids = np.asarray(query_ids, dtype=np.int64)
        for c0 in range(0, len(query_ids), SEARCH_CHUNK):
            nn_scores, nn_doc_ids = index_retrieve_arrays(index, query_embs[c0:c0 + SEARCH_CHUNK], args.top_k)
            stream.put(qids[c0:c0 + SEARCH_CHUNK], nn_doc_ids, nn_scores)
        print(f"{header} = {len(set(query_ids))}")
        avg = stream.close()
    print(f"average ranks per query = {avg}")


if __name__ == "__main__":
    main(get_args())
