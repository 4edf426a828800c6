from .losses import BPRLoss, DualLoss, ListwiseLoss, SampledSoftmaxLoss, create_loss_function

__all__ = ["BPRLoss", "DualLoss", "ListwiseLoss", "SampledSoftmaxLoss", "create_loss_function"]
