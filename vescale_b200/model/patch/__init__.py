"""Model patches for TP (legacy ``vescale/model/patch``)."""
from .linear import RowParallelLinear  # noqa: F401
from .vp_embedding import VocabParallelEmbedding  # noqa: F401
from .vp_cross_entropy import VocabParallelCrossEntropy  # noqa: F401
