"""Recall@K / NDCG@K with the reference's signatures (etpgt/utils/metrics.py), computed by the
device-side accumulator kernel; one host read for the final float."""

from __future__ import annotations

import torch

from .. import ops


def _rates(predictions: torch.Tensor, targets: torch.Tensor, k: int):
    if predictions.size(0) == 0:
        return float("nan"), float("nan")
    if not predictions.is_cuda:
        predictions, targets = predictions.cuda(), targets.cuda()
    acc = ops.topk_metrics(predictions, targets.to(predictions.device), min(k, predictions.size(1)))
    hits, gain = acc.tolist()
    return hits / predictions.size(0), gain / predictions.size(0)


def compute_recall_at_k(predictions: torch.Tensor, targets: torch.Tensor, k: int) -> float:
    return _rates(predictions, targets, k)[0]


def compute_ndcg_at_k(predictions: torch.Tensor, targets: torch.Tensor, k: int) -> float:
    return _rates(predictions, targets, k)[1]
