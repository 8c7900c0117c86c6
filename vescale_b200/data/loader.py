"""Rank-consistent batch sampling + pinned prefetch.

Contract kept from the reference's example loaders (``legacy/examples/llama2_4D_finetune/data_loader.py:32-64``): every rank draws
the SAME global batch of window start offsets (there: a replicated DTensor ``randint``; here: a counter-based generator keyed by
``(seed, split, step)`` — no communication, and a resumed run regenerates step ``k`` without replaying steps ``0..k-1``), then
data-parallel rank ``r`` keeps rows ``[r * local, (r + 1) * local)``; tensor-/pipeline-parallel peers of one DP rank therefore
see identical tokens.

B200-side of the pipeline: batches are assembled by a background thread straight into a ring of PINNED host buffers and copied
with ``non_blocking=True`` on a dedicated copy stream; ``next()`` makes the compute stream wait on the copy's event, so the H2D
of step ``k + 1`` overlaps the compute of step ``k`` and the training loop never blocks on the page cache.  On CPU (tests) the
same code runs without pinning and streams.
"""
from __future__ import annotations

import queue
import threading
from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from .dataset import TokenBinDataset

__all__ = ["DistributedTokenLoader", "sample_indices"]

_SPLIT_ID = {"train": 0, "val": 1, "test": 2}


def sample_indices(n_windows: int, batch: int, seed: int, step: int, split: str = "train") -> np.ndarray:
    """``batch`` window starts in ``[0, n_windows)``: a pure function of ``(seed, split, step)`` (Philox counter-based)."""
    bitgen = np.random.Philox(key=[int(seed) & (2**64 - 1), _SPLIT_ID.get(split, 3)], counter=[int(step), 0, 0, 0])
    return np.random.Generator(bitgen).integers(0, n_windows, size=batch, dtype=np.int64)


class DistributedTokenLoader:
    """``for x, y in loader`` / ``loader.get_batch(step)``: next-token-prediction batches ``x = tokens[i : i+S]``, ``y = tokens[i+1 : i+S+1]``.

    ``global_batch`` rows are sampled per step for the whole job; this rank returns its data-parallel slice.  ``dp_rank / dp_size``
    can be given directly or taken from ``mesh`` (dimension ``dp_dim``, default ``"DP"``)."""

    def __init__(self, dataset: TokenBinDataset, seq_len: int, global_batch: int, *, dp_rank: int = 0, dp_size: int = 1, mesh=None, dp_dim="DP",
                 device="cpu", seed: int = 1337, split: str = "train", prefetch: int = 2, start_step: int = 0, pin: Optional[bool] = None):
        if mesh is not None:
            names = getattr(mesh, "mesh_dim_names", None) or ()
            d = names.index(dp_dim) if isinstance(dp_dim, str) and dp_dim in names else (dp_dim if isinstance(dp_dim, int) else 0)
            dp_rank, dp_size = mesh.get_local_rank(d), mesh.size(d)
        if global_batch % dp_size:
            raise ValueError(f"global batch {global_batch} is not divisible by the data-parallel size {dp_size}")
        if len(dataset) < seq_len + 2:
            raise ValueError(f"dataset has {len(dataset)} tokens, need more than seq_len + 1 = {seq_len + 1}")
        self.ds, self.S, self.B, self.rank, self.size = dataset, seq_len, global_batch, dp_rank, dp_size
        self.local = global_batch // dp_size
        self.device = torch.device(device)
        self.seed, self.split = seed, split
        self.step = start_step
        self.cuda = self.device.type == "cuda"
        self.pin = self.cuda if pin is None else pin
        self.depth = max(1, prefetch)
        self._ring = [self._host_buf() for _ in range(self.depth + 1)]
        self._free: "queue.Queue[int]" = queue.Queue()
        for i in range(len(self._ring)):
            self._free.put(i)
        self._ready: "queue.Queue[Tuple[int, int]]" = queue.Queue()
        self._copy_stream = torch.cuda.Stream(self.device) if self.cuda else None
        self._thread: Optional[threading.Thread] = None
        self._stop = threading.Event()
        self._inflight = []  # (slot, event) of device copies whose host buffer is not reusable yet
        self.h2d_bytes_per_step = 2 * self.local * seq_len * 8

    # ------------------------------------------------------------------ plumbing
    def _host_buf(self) -> torch.Tensor:
        t = torch.empty(self.local, self.S + 1, dtype=torch.int64)
        return t.pin_memory() if self.pin and torch.cuda.is_available() else t

    def indices(self, step: int) -> np.ndarray:
        """This rank's window starts at ``step`` (rows ``[rank * local, (rank + 1) * local)`` of the global draw)."""
        ix = sample_indices(len(self.ds) - self.S - 1, self.B, self.seed, step, self.split)
        return ix[self.rank * self.local : (self.rank + 1) * self.local]

    def _assemble(self, slot: int, step: int) -> None:
        self.ds.fill(self._ring[slot], self.indices(step))

    def _worker(self, first_step: int) -> None:
        step = first_step
        while not self._stop.is_set():
            try:
                slot = self._free.get(timeout=0.1)
            except queue.Empty:
                continue
            self._assemble(slot, step)
            self._ready.put((slot, step))
            step += 1

    def _ensure_thread(self) -> None:
        if self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._worker, args=(self.step,), daemon=True, name="vescale-data")
            self._thread.start()

    def _recycle(self) -> None:
        keep = []
        for slot, ev in self._inflight:
            if ev is None or ev.query():
                self._free.put(slot)
            else:
                keep.append((slot, ev))
        self._inflight = keep

    # ------------------------------------------------------------------ API
    def get_batch(self, step: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """Synchronous form (evaluation, tests, or random access by ``step``): no thread, no ring."""
        st = self.step if step is None else step
        buf = torch.empty(self.local, self.S + 1, dtype=torch.int64)
        self.ds.fill(buf, self.indices(st))
        if step is None:
            self.step += 1
        dev = buf.to(self.device)
        return dev[:, :-1], dev[:, 1:]

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        return self

    def __next__(self) -> Tuple[torch.Tensor, torch.Tensor]:
        self._ensure_thread()
        self._recycle()
        slot, step = self._ready.get()
        assert step == self.step, (step, self.step)
        host = self._ring[slot]
        if self.cuda:
            with torch.cuda.stream(self._copy_stream):
                dev = host.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            torch.cuda.current_stream(self.device).wait_event(ev)
            dev.record_stream(torch.cuda.current_stream(self.device))
            self._inflight.append((slot, ev))
        else:
            dev = host.clone()
            self._free.put(slot)
        self.step += 1
        return dev[:, :-1], dev[:, 1:]

    def state_dict(self) -> dict:
        return {"step": self.step, "seed": self.seed, "split": self.split}

    def load_state_dict(self, sd: dict) -> None:
        """Resume at ``sd['step']``: the sampler is counter-based, so nothing is replayed."""
        self.close()
        self.step, self.seed, self.split = int(sd["step"]), int(sd.get("seed", self.seed)), sd.get("split", self.split)

    def close(self) -> None:
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=5)
            self._thread = None
        if self.cuda:
            torch.cuda.synchronize(self.device)
        while not self._ready.empty():
            self._ready.get()
        while not self._free.empty():
            self._free.get()
        self._inflight = []
        for i in range(len(self._ring)):
            self._free.put(i)

    def __del__(self):
        try:
            self._stop.set()
        except Exception:
            pass
