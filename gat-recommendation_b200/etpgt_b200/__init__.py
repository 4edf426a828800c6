"""etpgt_b200 — B200-native (sm_100a) implementation of the train/eval hot path of the `etpgt`
session recommender, behind the reference's own module API:

    from etpgt_b200.model import create_graph_transformer_optimized
    from etpgt_b200.train.losses import create_loss_function

Every op on the path runs in libetpgt_b200.so (C ABI in include/etpgt_b200.h); importing this
package without the built library raises ImportError — there is no CPU / PyTorch fallback.
"""

from . import _lib

_lib.load()  # fail loudly at import time if the CUDA library is missing

from . import ops  # noqa: E402
from .model import (  # noqa: E402
    BaseRecommendationModel,
    GraphTransformer,
    SessionReadout,
    create_graph_transformer,
    create_graph_transformer_optimized,
)
from .train.losses import create_loss_function  # noqa: E402

__version__ = "0.1.0"
__all__ = ["ops", "BaseRecommendationModel", "SessionReadout", "GraphTransformer", "create_graph_transformer",
           "create_graph_transformer_optimized", "create_loss_function"]
