"""Transposed search (passage -> top-k queries): the reference's retriever/retrieve_top_queries.py
(:27-82), whose `models.dual_encoder` import is broken upstream.  Same flags and run-file layout
`pid \\t qid \\t rank \\t score`; shares the implementation of retrieve_top_passages with roles swapped."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retriever import retrieve_top_passages as rtp  # noqa: E402


def get_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--resume", default=None)
    parser.add_argument("--model_name_or_path", default="sebastian-hofstaetter/distilbert-dot-tas_b-b256-msmarco")
    parser.add_argument("--tokenizer_name_or_path", default="distilbert-base-uncased")
    parser.add_argument("--passages_path", default="passages.dev.small.tsv")
    parser.add_argument("--index_path", default="")
    parser.add_argument("--max_length", default=256, type=int)
    parser.add_argument("--top_k", default=200, type=int)
    parser.add_argument("--output_path", default="")
    parser.add_argument("--gpus", default="0")
    parser.add_argument("--precision", default="auto", choices=["auto", "f16", "bf16", "tf32", "simt"])
    args = parser.parse_args(argv)
    # the reference builds DualEncoder(..., share_weights=True) and loads the state dict as it is
    args.share_weights = True
    args.is_parallel = False
    args.queries_path = args.passages_path
    return args


def main(args):
    # the reference has no path guards in this script
    rtp.main(args, is_query_side=False, header="# unique passages", guards=False)


if __name__ == "__main__":
    main(get_args())
