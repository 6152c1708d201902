"""The encoder side of the retrieval pipeline, kept as plain PyTorch / HF modules (the north star
keeps PyTorch for tensor plumbing; the encoder is a feeder, not a kernel target).

Mirrors models/nway_dual_encoder.py:51-57 (CLS vector of the last hidden state) and
dataset/sequence_dataset.py:41-66 (TSV parsing, first-seen order, padding to the longest)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch
from torch.utils.data import Dataset


class DualEncoder(torch.nn.Module):
    """query_embs / passage_embs -> [B, hidden] CLS vectors.  share_weights=False keeps two towers
    named query_encoder / passage_encoder like the reference, so its checkpoints load as they are."""

    def __init__(self, model_name_or_path: str, share_weights: bool = False):
        super().__init__()
        from transformers import AutoModel
        self.share_weights = share_weights
        self.query_encoder = AutoModel.from_pretrained(model_name_or_path)
        self.passage_encoder = self.query_encoder if share_weights else AutoModel.from_pretrained(model_name_or_path)

    def query_embs(self, query):
        return self.query_encoder(**query)[0][:, 0, :]

    def passage_embs(self, passage):
        return self.passage_encoder(**passage)[0][:, 0, :]


def load_checkpoint(model: torch.nn.Module, path: str, is_parallel: bool = True) -> None:
    """retriever/retrieve_top_passages.py:63-75: strip the DDP `module.` prefix when asked."""
    checkpoint = torch.load(path, map_location="cpu")
    state_dict = checkpoint["state_dict"]
    if is_parallel:
        print("load parallel wrapped model.")
        state_dict = OrderedDict((k[7:], v) for k, v in state_dict.items())  # remove `module.`
    model.load_state_dict(state_dict)


class SequenceDataset(Dataset):
    def __init__(self, id_to_seq: Dict[int, str], tokenizer, max_length: int, is_query: bool):
        self.tokenizer, self.max_length, self.is_query = tokenizer, int(max_length), is_query
        self.id_seq_pair = list(id_to_seq.items())

    def __getitem__(self, idx):
        sid, seq = self.id_seq_pair[idx]
        return {"seq": seq, "id": sid}

    def __len__(self):
        return len(self.id_seq_pair)

    @classmethod
    def create_from_seqs_file(cls, seqs_file, tokenizer, max_length, is_query):
        id_to_seq: Dict[int, str] = {}
        with open(seqs_file, "r") as f:
            for line in f:
                sid, seq = line.strip().split("\t")   # exactly two fields, like the reference
                id_to_seq[int(sid)] = seq             # duplicates: first position, last text
        return cls(id_to_seq, tokenizer, max_length, is_query)

    def collate_fn(self, batch):
        ids = [e["id"] for e in batch]
        seqs = self.tokenizer([e["seq"] for e in batch], padding=True, truncation="longest_first",
                              return_tensors="pt", max_length=self.max_length)
        return {"seq": seqs, "id": ids}
