from .base import BaseRecommendationModel, SessionReadout
from .graph_transformer import GraphTransformer, create_graph_transformer, create_graph_transformer_optimized

__all__ = ["BaseRecommendationModel", "SessionReadout", "GraphTransformer", "create_graph_transformer",
           "create_graph_transformer_optimized"]
