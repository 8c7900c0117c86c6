"""Token data pipeline (reference contract: ``legacy/examples/llama2_4D_finetune/data_loader.py`` — every rank draws the same global
batch of offsets, DP rank r keeps its slice; files are nanoGPT-style flat token ``.bin``)."""
import numpy as np
import torch

from vescale_b200.data import DistributedTokenLoader, TokenBinDataset, prepare_char_corpus, sample_indices, write_token_bin


def test_token_bin_dataset_and_prepare(tmp_path):
    meta = prepare_char_corpus(str(tmp_path / "char"), n_chars=5000)
    assert meta == prepare_char_corpus(str(tmp_path / "char"))  # idempotent
    ds = TokenBinDataset(str(tmp_path / "char" / "train.bin"))
    assert len(ds) == meta["train_tokens"] and int(ds.window(0, len(ds)).max()) < meta["vocab_size"]
    big = write_token_bin(str(tmp_path / "big.bin"), np.array([1, 70000, 128255]))
    d32 = TokenBinDataset(big, dtype=np.uint32)
    assert d32.window(0, 3).tolist() == [1, 70000, 128255]
    import pickle

    ds.data  # open the memmap, then make sure it is not pickled
    assert pickle.loads(pickle.dumps(ds))._mm is None


def test_loader_rank_consistency_prefetch_and_resume(tmp_path):
    toks = np.arange(10000) % 251
    ds = TokenBinDataset(write_token_bin(str(tmp_path / "t.bin"), toks))
    S, B = 16, 8
    full = DistributedTokenLoader(ds, S, B, seed=7)
    parts = [DistributedTokenLoader(ds, S, B, dp_rank=r, dp_size=4, seed=7) for r in range(4)]
    for step in range(3):
        x, y = full.get_batch(step)
        assert torch.equal(x[:, 1:], y[:, :-1])  # next-token targets
        xs = torch.cat([p.get_batch(step)[0] for p in parts])
        assert torch.equal(xs, x)  # the DP slices tile the global batch, in order
    assert not np.array_equal(sample_indices(1000, 8, 7, 0), sample_indices(1000, 8, 7, 1))
    assert not np.array_equal(sample_indices(1000, 8, 7, 0, "train"), sample_indices(1000, 8, 7, 0, "val"))
    # the prefetching iterator yields exactly the random-access batches, in step order
    it = DistributedTokenLoader(ds, S, B, dp_rank=1, dp_size=2, seed=7, prefetch=3)
    got = [next(it) for _ in range(6)]
    for step, (x, y) in enumerate(got):
        rx, ry = DistributedTokenLoader(ds, S, B, dp_rank=1, dp_size=2, seed=7).get_batch(step)
        assert torch.equal(x, rx) and torch.equal(y, ry)
    # resume: a fresh loader restored at step 4 continues with step 4 without replaying 0..3
    sd = {"step": 4, "seed": 7, "split": "train"}
    it.close()
    fresh = DistributedTokenLoader(ds, S, B, dp_rank=1, dp_size=2, seed=0)
    fresh.load_state_dict(sd)
    x4, _ = next(fresh)
    assert torch.equal(x4, got[4][0]) and fresh.state_dict()["step"] == 5
    fresh.close()
    try:
        DistributedTokenLoader(ds, S, 6, dp_size=4)
        raise AssertionError("indivisible global batch must be rejected")
    except ValueError:
        pass
