"""TEST INFRASTRUCTURE ONLY — CPU oracle for the model-level hot path (SURVEY.md §8a rows
a4-a10, a12), written as pure functions of a reference-format `state_dict`.

Each function restates one piece of the reference and cites it.  The restatement is pinned
against the UNMODIFIED reference classes (run on the `oracle/pyg_shim` stand-in for the absent
`torch_geometric`) by `oracle/make_golden.py` -> `tests/golden/*.npz` and
`tests/test_oracle.py`.  Below the PyG boundary parity is unpinned (see `conv_ref.py`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module; it is the checker, never the product.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import conv_ref


@dataclass
class BatchNormOut:
    y: torch.Tensor
    running_mean: torch.Tensor
    running_var: torch.Tensor


def embed_with_pe(state: dict, ids: torch.Tensor, pe_rows: torch.Tensor | None = None) -> torch.Tensor:
    """x0 = E[ids] (+ pe[ids] @ W_pe^T + b_pe).
    etpgt/model/graph_transformer.py:140-152, etpgt/encodings/laplacian_pe.py:170-199."""
    x = state["item_embedding.weight"][ids]
    if "laplacian_pe.projection.weight" in state:
        if pe_rows is None:
            cached = state.get("laplacian_pe._cached_pe")
            if cached is None:
                raise RuntimeError("Laplacian PE not precomputed. Call precompute() first.")
            pe_rows = cached[ids]
        x = x + F.linear(pe_rows, state["laplacian_pe.projection.weight"], state["laplacian_pe.projection.bias"])
    return x


def batch_norm_rows(x, weight, bias, running_mean, running_var, training, momentum=0.1, eps=1e-5) -> BatchNormOut:
    """nn.BatchNorm1d over the N node rows of the whole batch
    (etpgt/model/graph_transformer.py:83,175): biased variance normalises, the unbiased one
    feeds the running statistic."""
    if training:
        n = x.size(0)
        mean = x.mean(dim=0)
        var = x.var(dim=0, unbiased=False)
        new_mean = (1.0 - momentum) * running_mean + momentum * mean.detach()
        new_var = (1.0 - momentum) * running_var + momentum * var.detach() * (n / max(n - 1, 1))
    else:
        mean, var, new_mean, new_var = running_mean, running_var, running_mean, running_var
    y = (x - mean) / torch.sqrt(var + eps) * weight + bias
    return BatchNormOut(y, new_mean, new_var)


def session_readout(x, batch_vec, num_sessions, kind="mean", att_w=None, att_b=None) -> torch.Tensor:
    """etpgt/model/base.py:136-193.  Sessions are contiguous node ranges (PyG collate), so
    "last" is the highest-index node of the range."""
    rows = []
    for s in range(num_sessions):
        seg = x[batch_vec == s]
        if kind == "mean":
            rows.append(seg.mean(dim=0))
        elif kind == "max":
            rows.append(seg.max(dim=0)[0])
        elif kind == "last":
            rows.append(seg[-1])
        elif kind == "attention":
            score = F.linear(seg, att_w, att_b).squeeze(-1)
            rows.append(torch.softmax(score, dim=0) @ seg)
        else:
            raise ValueError(f"Unknown readout type: {kind}")
    return torch.stack(rows)


def graph_transformer_forward(state, ids, edge_index, batch_vec, *, num_layers, num_heads,
                              readout="mean", training=False, pe_rows=None, use_ffn=False,
                              return_nodes=False):
    """etpgt/model/graph_transformer.py:126-182 with dropout p=0 (parity runs use dropout 0
    or eval mode; SURVEY.md §7 'Hard parts')."""
    x = embed_with_pe(state, ids, pe_rows)
    new_stats = {}
    for layer in range(num_layers):
        p = f"convs.{layer}."
        res = x
        x = conv_ref.transformer_conv(
            x, edge_index,
            state[p + "lin_query.weight"], state[p + "lin_query.bias"],
            state[p + "lin_key.weight"], state[p + "lin_key.bias"],
            state[p + "lin_value.weight"], state[p + "lin_value.bias"],
            state[p + "lin_skip.weight"], state[p + "lin_skip.bias"],
            state[p + "lin_beta.weight"], num_heads)
        b = f"batch_norms.{layer}."
        bn = batch_norm_rows(x, state[b + "weight"], state[b + "bias"], state[b + "running_mean"],
                             state[b + "running_var"], training)
        new_stats[b + "running_mean"], new_stats[b + "running_var"] = bn.running_mean, bn.running_var
        x = bn.y + res
        if use_ffn:
            f = f"ffns.{layer}."
            h = F.gelu(F.linear(x, state[f + "0.weight"], state[f + "0.bias"]))
            x = x + F.linear(h, state[f + "3.weight"], state[f + "3.bias"])
    num_sessions = int(batch_vec.max()) + 1
    out = session_readout(x, batch_vec, num_sessions, readout,
                          state.get("readout.attention.weight"), state.get("readout.attention.bias"))
    return (out, x, new_stats) if return_nodes else out


def gat_forward(state, ids, edge_index, batch_vec, *, num_convs, num_heads, readout="mean",
                training=False, concat_heads=False):
    """etpgt/model/gat.py:119-146 (dropout 0)."""
    x = state["item_embedding.weight"][ids]
    for layer in range(num_convs):
        p = f"convs.{layer}."
        concat = concat_heads and layer < num_convs - 1
        x = conv_ref.gat_conv(x, edge_index, state[p + "lin.weight"], state[p + "att_src"],
                              state[p + "att_dst"], state[p + "bias"], num_heads, concat)
        b = f"batch_norms.{layer}."
        x = batch_norm_rows(x, state[b + "weight"], state[b + "bias"], state[b + "running_mean"],
                            state[b + "running_var"], training).y
        if layer < num_convs - 1:
            x = torch.relu(x)
    return session_readout(x, batch_vec, int(batch_vec.max()) + 1, readout,
                           state.get("readout.attention.weight"), state.get("readout.attention.bias"))


def graphsage_forward(state, ids, edge_index, batch_vec, *, num_layers, readout="mean", training=False):
    """etpgt/model/graphsage.py:57-83 (dropout 0)."""
    x = state["item_embedding.weight"][ids]
    for layer in range(num_layers):
        p = f"convs.{layer}."
        x = conv_ref.sage_conv(x, edge_index, state[p + "lin_l.weight"], state[p + "lin_l.bias"],
                               state[p + "lin_r.weight"])
        b = f"batch_norms.{layer}."
        x = torch.relu(batch_norm_rows(x, state[b + "weight"], state[b + "bias"], state[b + "running_mean"],
                                       state[b + "running_var"], training).y)
    return session_readout(x, batch_vec, int(batch_vec.max()) + 1, readout,
                           state.get("readout.attention.weight"), state.get("readout.attention.bias"))


# ----------------------------------------------------------------------------- losses


def sampled_scores(sess, table, targets, negatives):
    """pos = <S, E[t]>, neg = <S, E[n]>  (etpgt/train/losses.py:39-48, etpgt/model/base.py:97-108)."""
    pos = (sess * table[targets]).sum(dim=1)
    neg = torch.einsum("bnd,bd->bn", table[negatives], sess)
    return pos, neg


def bpr_loss(sess, table, targets, negatives):
    """-log(sigmoid(pos - neg) + 1e-8).mean() over B*neg (etpgt/train/losses.py:51)."""
    pos, neg = sampled_scores(sess, table, targets, negatives)
    return -torch.log(torch.sigmoid(pos.unsqueeze(1) - neg) + 1e-8).mean()


def listwise_loss(sess, table, targets, negatives, temperature=1.0):
    """Cross-entropy of [pos, negs] / T against index 0 (etpgt/train/losses.py:98-109)."""
    pos, neg = sampled_scores(sess, table, targets, negatives)
    logits = torch.cat([pos.unsqueeze(1), neg], dim=1) / temperature
    return (torch.logsumexp(logits, dim=1) - logits[:, 0]).mean()


def dual_loss(sess, table, targets, negatives, alpha=0.7, temperature=1.0):
    """alpha * listwise + (1 - alpha) * bpr (etpgt/train/losses.py:149-155)."""
    lw = listwise_loss(sess, table, targets, negatives, temperature)
    bp = bpr_loss(sess, table, targets, negatives)
    return alpha * lw + (1.0 - alpha) * bp, lw, bp


# ----------------------------------------------------------------------------- scoring


def topk_lower_id(scores: torch.Tensor, k: int):
    """Top-k by score descending, ties to the lower item id.  `torch.topk` (base.py:76) leaves
    ties unspecified; BASELINE.json defines them, and a stable descending sort realises it."""
    order = torch.sort(scores, dim=1, descending=True, stable=True)
    return order.values[:, :k], order.indices[:, :k]


def predict(sess, table, k=20, bf16_inputs=False):
    """scores = S @ E^T, top-k ids (etpgt/model/base.py:59-78).  With `bf16_inputs` both
    operands are rounded to bf16 first and accumulated in fp32/fp64 — the oracle for the
    tensor-core scoring kernel."""
    if bf16_inputs:
        sess = sess.to(torch.bfloat16).to(torch.float64)
        table = table.to(torch.bfloat16).to(torch.float64)
    return topk_lower_id(sess @ table.t(), k)


def recall_at_k(topk_ids, targets, k):
    """etpgt/utils/metrics.py:21-27."""
    return (topk_ids[:, :k] == targets.unsqueeze(1)).any(dim=1).double().mean().item()


def ndcg_at_k(topk_ids, targets, k):
    """etpgt/utils/metrics.py:45-66: 1/log2(pos+2) for the single relevant item."""
    hit = topk_ids[:, :k] == targets.unsqueeze(1)
    pos = hit.double().argmax(dim=1)
    gain = torch.where(hit.any(dim=1), 1.0 / torch.log2(pos.double() + 2.0), torch.zeros_like(pos, dtype=torch.float64))
    return gain.mean().item()


def scale(c: int) -> float:
    return 1.0 / math.sqrt(c)
