"""Token files on disk.

The reference's examples read ``data/<dataset>/{train,val}.bin`` — flat ``uint16`` token streams written by a ``prepare.py``
(nanoGPT's convention) — through ``np.memmap`` (``legacy/examples/llama2_4D_finetune/data_loader.py:32-64``,
``legacy/examples/nanogpt_4D_finetune/finetune_4D.py``).  Same file format here (so files prepared for the reference load
unchanged), plus a small header-less ``uint32`` variant for vocabularies above 65 535 (Llama-3's 128 256) chosen by file size
hint or ``dtype=``.

There is no network in the build environment, so ``prepare_char_corpus`` can also write a *synthetic but learnable* corpus
(sentences from a small grammar) to stand in for Shakespeare; character-level encoding is the reference's
``shakespeare_char`` recipe.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

__all__ = ["TokenBinDataset", "write_token_bin", "encode_chars", "prepare_char_corpus", "synthetic_corpus"]


def write_token_bin(path: str, tokens, dtype=None) -> str:
    """Write a flat token stream.  ``dtype`` defaults to ``uint16`` when every id fits, else ``uint32``."""
    arr = np.asarray(tokens.cpu().numpy() if isinstance(tokens, torch.Tensor) else tokens)
    if dtype is None:
        dtype = np.uint16 if (arr.size == 0 or int(arr.max()) < 2**16) else np.uint32
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    arr.astype(dtype).tofile(path)
    return path


class TokenBinDataset:
    """A memory-mapped token stream; ``window(i, n)`` is tokens ``[i, i + n)`` as an int64 tensor.

    The memmap is re-opened lazily per process (it must not be pickled into loader threads / forked workers with an open
    handle) and dropped by ``close()``; reads go through the page cache, so a hot dataset costs no copies beyond the one into
    the pinned batch buffer."""

    def __init__(self, path: str, dtype=None):
        self.path = path
        if dtype is None:
            meta = os.path.join(os.path.dirname(path), "meta.json")
            dtype = np.uint16
            if os.path.exists(meta):
                with open(meta) as f:
                    dtype = np.dtype(json.load(f).get("dtype", "uint16")).type
        self.dtype = np.dtype(dtype)
        size = os.path.getsize(path)
        if size % self.dtype.itemsize:
            raise ValueError(f"{path}: {size} bytes is not a whole number of {self.dtype} tokens")
        self.n_tokens = size // self.dtype.itemsize
        self._mm: Optional[np.memmap] = None

    def __len__(self) -> int:
        return self.n_tokens

    @property
    def data(self) -> np.memmap:
        if self._mm is None:
            self._mm = np.memmap(self.path, dtype=self.dtype, mode="r")
        return self._mm

    def close(self) -> None:
        self._mm = None

    def window(self, start: int, n: int) -> torch.Tensor:
        return torch.from_numpy(self.data[start : start + n].astype(np.int64))

    def fill(self, out: torch.Tensor, starts) -> torch.Tensor:
        """``out[b] = tokens[starts[b] : starts[b] + out.shape[1]]`` written in place (``out``: int64, typically pinned)."""
        n = out.shape[1]
        view = out.numpy()
        d = self.data
        for b, s in enumerate(starts):
            view[b, :] = d[int(s) : int(s) + n]
        return out

    def __getstate__(self):
        st = dict(self.__dict__)
        st["_mm"] = None
        return st


def encode_chars(text: str, stoi: Optional[Dict[str, int]] = None) -> Tuple[np.ndarray, Dict[str, int]]:
    """Character-level encoding (the ``shakespeare_char`` recipe): vocabulary = sorted set of characters."""
    if stoi is None:
        stoi = {ch: i for i, ch in enumerate(sorted(set(text)))}
    return np.array([stoi[c] for c in text], dtype=np.uint16), stoi


def synthetic_corpus(n_chars: int = 200_000, seed: int = 0) -> str:
    """Sentences from a small grammar with agreement between clauses: a stream a small LM can learn (loss falls well below the
    unigram entropy), unlike uniform noise.  Deterministic in ``seed``."""
    g = np.random.default_rng(seed)
    subj = ["the king", "my lord", "a fool", "thy brother", "the night", "sweet love", "old time", "the crown", "her ghost", "our house"]
    verb = ["doth speak", "shall fall", "will rise", "hath seen", "must die", "may weep", "did swear", "can wait"]
    tail = ["of war", "in sorrow", "to the sea", "with grace", "by night", "for gold", "no more", "at dawn"]
    link = [", and", ", yet", "; so", ", for"]
    parts = []
    total = 0
    while total < n_chars:
        i, j, k = int(g.integers(len(subj))), int(g.integers(len(verb))), int(g.integers(len(tail)))
        s = f"{subj[i]} {verb[j]} {tail[k]}"
        if g.random() < 0.5:  # a second clause whose verb repeats the first one's index (long-range structure)
            s += f"{link[int(g.integers(len(link)))]} {subj[int(g.integers(len(subj)))]} {verb[j]} {tail[int(g.integers(len(tail)))]}"
        s = s[0].upper() + s[1:] + ".\n"
        parts.append(s)
        total += len(s)
    return "".join(parts)[:n_chars]


def prepare_char_corpus(out_dir: str, text: Optional[str] = None, val_fraction: float = 0.1, n_chars: int = 200_000, seed: int = 0) -> Dict:
    """Write ``train.bin`` / ``val.bin`` / ``meta.json`` (vocabulary, dtype) under ``out_dir`` from ``text`` (a file's contents) or the
    synthetic corpus.  Idempotent: an existing prepared directory is returned as is."""
    meta_p = os.path.join(out_dir, "meta.json")
    if os.path.exists(meta_p) and os.path.exists(os.path.join(out_dir, "train.bin")):
        with open(meta_p) as f:
            return json.load(f)
    text = text if text is not None else synthetic_corpus(n_chars, seed)
    ids, stoi = encode_chars(text)
    n_val = max(1, int(len(ids) * val_fraction))
    write_token_bin(os.path.join(out_dir, "train.bin"), ids[:-n_val], np.uint16)
    write_token_bin(os.path.join(out_dir, "val.bin"), ids[-n_val:], np.uint16)
    meta = {"vocab_size": len(stoi), "stoi": stoi, "dtype": "uint16", "train_tokens": int(len(ids) - n_val), "val_tokens": int(n_val)}
    tmp = meta_p + f".tmp{os.getpid()}"
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, meta_p)
    return meta
