"""Model families of the reference (etpgt/model/__init__.py) on the B200 kernels."""

from .base import BaseRecommendationModel, SessionReadout
from .gat import GAT, create_gat
from .graph_transformer import GraphTransformer, create_graph_transformer, create_graph_transformer_optimized
from .graphsage import GraphSAGE, create_graphsage

__all__ = ["BaseRecommendationModel", "SessionReadout", "GraphSAGE", "create_graphsage", "GAT", "create_gat",
           "GraphTransformer", "create_graph_transformer", "create_graph_transformer_optimized"]
