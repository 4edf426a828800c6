"""`Trainer` with the reference's constructor, attributes and methods (etpgt/train/trainer.py:17-251:
`train_epoch() -> float`, `evaluate() -> {"recall@k", "ndcg@k"}`, `save_checkpoint(is_best)`, `train() ->
history`, early stopping on recall@k_values[0], `checkpoint_latest.pt / checkpoint_best.pt / history.json`).

What differs is where the work runs:
  * a training step goes through the C++ step driver (`FusedTrainStep`) when the model / loss pair is one it
    covers, else through the per-operator autograd path — both give the same bits;
  * the loss is accumulated on the device and read once per epoch (the reference syncs twice per step,
    trainer.py:130-133);
  * evaluation keeps predictions on the device: fused scoring + top-k, hit / NDCG counters accumulated by
    `etpgt_topk_metrics`, one read-back per evaluation (the reference copies every batch's top-k to the host,
    trainer.py:158-159).
"""

from __future__ import annotations

import json
import logging
from pathlib import Path

import torch

from .. import ops
from .losses import BPRLoss, DualLoss, ListwiseLoss
from .step import FusedTrainStep

logger = logging.getLogger(__name__)


def _driver_for(model, loss_fn):
    """FusedTrainStep for (model, loss_fn), or None when the pair needs the per-operator path."""
    if not FusedTrainStep.supported(model):
        return None
    if loss_fn is None or type(loss_fn) is BPRLoss:
        return FusedTrainStep(model, "bpr")
    if isinstance(loss_fn, ListwiseLoss):       # SampledSoftmaxLoss is the same arithmetic
        return FusedTrainStep(model, "listwise", temperature=loss_fn.temperature)
    if type(loss_fn) is DualLoss:
        return FusedTrainStep(model, "dual", alpha=loss_fn.alpha, temperature=loss_fn.listwise_loss.temperature)
    return None


class Trainer:
    def __init__(self, model, train_loader, val_loader, optimizer, device: str = "cuda",
                 output_dir: Path | str = "outputs", max_epochs: int = 100, patience: int = 10, eval_every: int = 1,
                 k_values: list[int] | None = None, loss_fn=None):
        self.model = model.to(device)
        self.train_loader, self.val_loader, self.optimizer = train_loader, val_loader, optimizer
        self.device = device
        self.output_dir = Path(output_dir)
        self.max_epochs, self.patience, self.eval_every = max_epochs, patience, eval_every
        self.k_values = k_values if k_values is not None else [10, 20]
        self.loss_fn = loss_fn
        self.output_dir.mkdir(parents=True, exist_ok=True)
        self.current_epoch = 0
        self.best_val_metric = 0.0
        self.patience_counter = 0
        self.history = {"train_loss": [], "val_metrics": []}
        self._driver = _driver_for(self.model, loss_fn)

    # ------------------------------------------------------------------ one step
    def _loss_and_backward(self, batch) -> torch.Tensor:
        batch_size = batch.target_item.shape[0]
        negatives = batch.negative_items.view(batch_size, -1)        # trainer.py:86-89
        if self._driver is not None and getattr(batch, "laplacian_pe", None) is None:
            return self._driver(batch, batch.target_item, negatives)[0]
        sess = self.model(batch)
        if self.loss_fn is not None:
            out = self.loss_fn(sess, batch.target_item, negatives, self.model.item_embedding)
            loss = out[0] if isinstance(out, tuple) else out         # DualLoss returns (loss, parts)
        else:
            loss = self.model.compute_loss(sess, batch.target_item, negatives)
        loss.backward()
        return loss.detach()

    def train_epoch(self) -> float:
        self.model.train()
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        steps = 0
        for batch in self.train_loader:
            batch = batch.to(self.device)
            self.optimizer.zero_grad()
            loss = self._loss_and_backward(batch)
            self.optimizer.step()
            total += loss
            steps += 1
        return float(total.item()) / max(steps, 1)

    # ------------------------------------------------------------------ evaluation
    @torch.no_grad()
    def evaluate(self) -> dict:
        self.model.eval()
        k_max = max(self.k_values)
        acc = {k: torch.zeros(2, dtype=torch.float64, device=self.device) for k in self.k_values}
        sessions = 0
        for batch in self.val_loader:
            batch = batch.to(self.device)
            top = self.model.predict(self.model(batch), k=k_max)
            for k in self.k_values:
                ops.topk_metrics(top, batch.target_item, k, acc[k])
            sessions += int(batch.target_item.shape[0])
        metrics = {}
        for k in self.k_values:
            hits, ndcg = acc[k].tolist()
            metrics[f"recall@{k}"] = hits / sessions if sessions else 0.0
            metrics[f"ndcg@{k}"] = ndcg / sessions if sessions else 0.0
        return metrics

    # ------------------------------------------------------------------ checkpoints and the outer loop
    def save_checkpoint(self, is_best: bool = False) -> None:
        checkpoint = {"epoch": self.current_epoch, "model_state_dict": self.model.state_dict(),
                      "optimizer_state_dict": self.optimizer.state_dict(),
                      "best_val_metric": self.best_val_metric, "history": self.history}
        names = ["checkpoint_latest.pt"] + (["checkpoint_best.pt"] if is_best else [])
        for name in names:
            torch.save(checkpoint, self.output_dir / name)
            logger.info("Saved checkpoint to %s", self.output_dir / name)

    def train(self) -> dict:
        logger.info("Starting training for %d epochs on %s", self.max_epochs, self.device)
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            train_loss = self.train_epoch()
            self.history["train_loss"].append(train_loss)
            logger.info("Epoch %d: train_loss=%.4f", epoch, train_loss)
            if (epoch + 1) % self.eval_every:
                continue
            metrics = self.evaluate()
            self.history["val_metrics"].append(metrics)
            logger.info("Epoch %d: %s", epoch, ", ".join(f"{k}={v:.4f}" for k, v in metrics.items()))
            watched = metrics[f"recall@{self.k_values[0]}"]
            is_best = watched > self.best_val_metric
            if is_best:
                self.best_val_metric, self.patience_counter = watched, 0
            else:
                self.patience_counter += 1
            self.save_checkpoint(is_best=is_best)
            if self.patience_counter >= self.patience:
                logger.info("Early stopping at epoch %d", epoch)
                break
        with open(self.output_dir / "history.json", "w") as f:
            json.dump(self.history, f, indent=2)
        return self.history
