"""Batched prediction driver and test metrics: the step right after the hot path.

``batch_posterior_computation`` mirrors ``cggp/cli_utils.py:426-436`` (a Python loop over batches of ``predict_fn``),
``test_metrics`` the metrics callback of ``cggp/optimize.py:285-364`` (``train/elbo`` is the caller's ``model.elbo``;
``test/rmse`` and ``test/nlpd`` here).  Unlike the reference nothing is pulled to the host per batch: means, variances
and the two reductions stay on the device, one scalar pair is read at the end.  With the test points sharded over
ranks (independent, no collective on the data path) the two sums are all-reduced once."""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import torch

from . import _lib


def batch_posterior_computation(predict_fn: Callable, data, batch_size: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """cli_utils.py:426-436: concatenated ``(means [N, 1], variances [N, 1])`` of ``predict_fn`` over row batches of
    ``data = (x, y)`` (``y`` is ignored, as in the reference)."""
    x = _lib.as_device_tensor(data[0])
    means, variances = [], []
    for s in range(0, x.shape[0], int(batch_size)):
        mean, variance = predict_fn(x[s:s + int(batch_size)])
        means.append(mean)
        variances.append(variance)
    if not means:
        empty = torch.empty((0, 1), dtype=x.dtype, device=x.device)
        return empty, empty.clone()
    return torch.cat(means, 0), torch.cat(variances, 0)


def test_metrics(model, test_data, batch_size: int, all_reduce: bool = False) -> Dict[str, float]:
    """optimize.py:300-350: ``rmse = sqrt(mean((y - mu)^2))``, ``nlpd = -sum(predict_log_density) / n`` with the
    model's likelihood, accumulated batch by batch on the device.  ``all_reduce=True``: ``test_data`` is this rank's
    shard of the test set and the sums are combined over the communicator of the context."""
    x = _lib.as_device_tensor(test_data[0])
    y = _lib.as_device_tensor(test_data[1], x.dtype)
    acc = torch.zeros(3, dtype=torch.float64, device=x.device)  # sum err^2, sum lpd, n
    for s in range(0, x.shape[0], int(batch_size)):
        xb, yb = x[s:s + int(batch_size)], y[s:s + int(batch_size)]
        mu, var = model.predict_f(xb)
        lpd = model.likelihood.predict_log_density(xb, mu, var, yb)
        err = yb - mu
        acc[0] += (err * err).sum().double()
        acc[1] += lpd.sum().double()
        acc[2] += err.numel()
    if all_reduce:
        _lib.context(x.device).allreduce_sum_(acc)
    sq, lpd, n = (float(v) for v in acc.cpu())
    n = max(n, 1.0)
    return {"test/rmse": (sq / n) ** 0.5, "test/nlpd": -lpd / n}


test_metrics.__test__ = False  # not a pytest test
