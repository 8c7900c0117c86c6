"""Model patches for TP (legacy ``vescale/model/patch``)."""
from .linear import RowParallelLinear  # noqa: F401
from .vp_embedding import VocabParallelEmbedding  # noqa: F401
from .vp_cross_entropy import VocabParallelCrossEntropy  # noqa: F401


def get_all_model_patch():
    """The patch entry points in application order (legacy ``model/patch/__init__.py:23``)."""
    return [RowParallelLinear.patch, VocabParallelEmbedding.patch, VocabParallelCrossEntropy.patch]
