"""Masked regression metrics with the semantics of the reference's ``Utils/util.py:510-559``
(mask = labels != null_val, renormalised by its mean; NaNs from empty masks become 0)."""
import math

import torch


def _mask(labels, null_val):
    if isinstance(null_val, float) and math.isnan(null_val):
        mask = ~torch.isnan(labels)
    else:
        mask = labels != null_val
    mask = mask.float()
    mask = mask / torch.mean(mask)
    return torch.where(torch.isnan(mask), torch.zeros_like(mask), mask)


def _reduce(loss, mask):
    loss = loss * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return torch.mean(loss)


def masked_mse(preds, labels, null_val=float("nan")):
    return _reduce((preds - labels) ** 2, _mask(labels, null_val))


def masked_rmse(preds, labels, null_val=float("nan")):
    return torch.sqrt(masked_mse(preds, labels, null_val))


def masked_mae(preds, labels, null_val=float("nan")):
    return _reduce(torch.abs(preds - labels), _mask(labels, null_val))


def masked_mape(preds, labels, null_val=float("nan")):
    return _reduce(torch.abs(preds - labels) / labels, _mask(labels, null_val))


def metric(pred, real):
    return (masked_mae(pred, real, 0.0).item(), masked_mape(pred, real, 0.0).item(), masked_rmse(pred, real, 0.0).item())


class StandardScaler:
    """Utils/util.py:104-117."""

    def __init__(self, mean, std):
        self.mean = mean
        self.std = std

    def transform(self, data):
        return (data - self.mean) / self.std

    def inverse_transform(self, data):
        return (data * self.std) + self.mean
