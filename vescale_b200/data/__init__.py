"""Token data pipeline for the training examples: ``.bin`` token files (memory-mapped), rank-consistent sampling, pinned-memory
prefetch with the host->device copy on a side stream."""
from .dataset import TokenBinDataset, encode_chars, prepare_char_corpus, synthetic_corpus, write_token_bin  # noqa: F401
from .loader import DistributedTokenLoader, sample_indices  # noqa: F401
