"""GraphTransformer (standard and optimized) on the B200 kernels; constructor arguments,
attributes and state-dict keys follow etpgt/model/graph_transformer.py."""

from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..encodings.laplacian_pe import LaplacianPECached
from ..nn import TransformerConv, batch_norm_rows, fused_layer_supported, transformer_layer
from .base import BaseRecommendationModel, SessionReadout


class GraphTransformer(BaseRecommendationModel):
    def __init__(self, num_items: int, embedding_dim: int = 256, hidden_dim: int = 256, num_layers: int = 3,
                 num_heads: int = 4, dropout: float = 0.1, readout_type: str = "mean",
                 use_laplacian_pe: bool = True, laplacian_k: int = 16, use_ffn: bool = True,
                 ffn_expansion: int = 4):
        super().__init__(num_items, embedding_dim, hidden_dim, num_layers, dropout)
        self.num_heads, self.readout_type = num_heads, readout_type
        self.use_laplacian_pe, self.laplacian_k = use_laplacian_pe, laplacian_k
        self.use_ffn, self.ffn_expansion = use_ffn, ffn_expansion
        if use_laplacian_pe:
            self.laplacian_pe = LaplacianPECached(k=laplacian_k, embedding_dim=embedding_dim)
        widths = [embedding_dim] + [hidden_dim] * (num_layers - 1)
        self.convs = nn.ModuleList(
            TransformerConv(w, hidden_dim // num_heads, heads=num_heads, dropout=dropout, concat=True, beta=True)
            for w in widths)
        self.batch_norms = nn.ModuleList(nn.BatchNorm1d(hidden_dim) for _ in widths)
        self.ffns = nn.ModuleList(self._make_ffn(hidden_dim) for _ in widths) if use_ffn else None
        self.dropout_layer = nn.Dropout(dropout)
        self.readout = SessionReadout(hidden_dim, readout_type)

    def _make_ffn(self, hidden_dim: int) -> nn.Module:
        inner = hidden_dim * self.ffn_expansion
        return nn.Sequential(nn.Linear(hidden_dim, inner), nn.GELU(), nn.Dropout(self.dropout),
                             nn.Linear(inner, hidden_dim), nn.Dropout(self.dropout))

    def _ffn_block(self, ffn: nn.Sequential, x):
        """x + Linear -> GELU -> Dropout -> Linear -> Dropout (graph_transformer.py:109-124,163-168): GEMM 1 with the
        GELU (+ Philox dropout) in its epilogue feeding GEMM 2 through split-bf16 operands, residual added by the
        second GEMM's TMA reduce (ops.FeedForward); the module (and its state-dict keys ffns.{l}.{0,3}) is the
        reference's nn.Sequential."""
        lin1, act, drop1, lin2, drop2 = ffn
        if ops.feed_forward_supported(x, lin1.weight, lin2.weight) and lin1.bias is not None and lin2.bias is not None:
            p1 = drop1.p if self.training else 0.0
            p2 = drop2.p if self.training else 0.0
            seeds = torch.randint(0, 2 ** 62, (2,)).tolist() if (p1 > 0.0 or p2 > 0.0) else (0, 0)
            return ops.FeedForward.apply(x, lin1.weight, lin1.bias, lin2.weight, lin2.bias, p1, p2, seeds[0], seeds[1])
        h = drop1(act(ops.linear(x, lin1.weight, lin1.bias)))
        return x + drop2(ops.linear(h, lin2.weight, lin2.bias))

    def forward(self, batch):
        ids, index = self._graph(batch)
        pe = w_pe = b_pe = None
        per_node = False
        if self.use_laplacian_pe:
            given = getattr(batch, "laplacian_pe", None)
            pe, per_node = (given, True) if given is not None else (self.laplacian_pe.cached(), False)
            w_pe, b_pe = self.laplacian_pe.projection.weight, self.laplacian_pe.projection.bias
        # bf16 hi/lo split of x for the fused layer's GEMM: written by the embedding kernel for the first layer
        # and by the previous fused layer's BatchNorm apply after that
        split = None
        if ids.numel() > 0 and ops.embed_split_supported(self.item_embedding.weight, self.embedding_dim,
                                                         self.hidden_dim):
            x, hi, lo = ops.EmbedPE.apply(ids, self.item_embedding.weight, pe, per_node, w_pe, b_pe,
                                          self.item_embedding.padding_idx, True)
            split = (hi, lo)
        else:
            x = ops.EmbedPE.apply(ids, self.item_embedding.weight, pe, per_node, w_pe, b_pe,
                                  self.item_embedding.padding_idx)
        for layer, (conv, bn) in enumerate(zip(self.convs, self.batch_norms)):
            if fused_layer_supported(conv, bn, x):
                drop_p = self.dropout_layer.p if self.training else 0.0
                want_split = not self.use_ffn and layer + 1 < len(self.convs)
                x, split = transformer_layer(conv, bn, x, split, index, drop_p, self.bn_process_group, want_split)
            else:
                x = batch_norm_rows(bn, conv(x, index), residual=x, group=self.bn_process_group)
                x = self.dropout_layer(x)
                split = None
            if self.use_ffn:
                x = self._ffn_block(self.ffns[layer], x)
        return self.readout(x, batch.batch, self._num_sessions(batch))


def create_graph_transformer(num_items: int, embedding_dim: int = 256, hidden_dim: int = 256, num_layers: int = 3,
                             num_heads: int = 4, dropout: float = 0.1, readout_type: str = "mean",
                             use_laplacian_pe: bool = True, laplacian_k: int = 16, use_ffn: bool = True,
                             ffn_expansion: int = 4) -> GraphTransformer:
    return GraphTransformer(num_items, embedding_dim, hidden_dim, num_layers, num_heads, dropout, readout_type,
                            use_laplacian_pe, laplacian_k, use_ffn, ffn_expansion)


def create_graph_transformer_optimized(num_items: int, embedding_dim: int = 256, hidden_dim: int = 256,
                                       num_layers: int = 2, num_heads: int = 2, dropout: float = 0.1,
                                       readout_type: str = "mean", use_laplacian_pe: bool = True,
                                       laplacian_k: int = 16, use_ffn: bool = False,
                                       ffn_expansion: int = 2) -> GraphTransformer:
    """The production configuration: 2 layers, 2 heads, no FFN (graph_transformer.py:231-280)."""
    return GraphTransformer(num_items, embedding_dim, hidden_dim, num_layers, num_heads, dropout, readout_type,
                            use_laplacian_pe, laplacian_k, use_ffn, ffn_expansion)
