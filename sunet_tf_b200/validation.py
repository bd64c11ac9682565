"""Validation-loop forward of train.py:420-470 over the fused eval epilogue (sunet_forward_eval).

Per batch the reference runs ``logits = model(input)``, ``prob = sigmoid(logits)``, ``se = (logits - target) ** 2`` and
three reductions (mean se, weighted se, weighted Charbonnier), each forcing a host sync through ``.item()``.  Here the
reductions are accumulated in float64 by the last kernel of the forward and read back once per batch (or once per epoch).
The morphological weight map (make_weights_from_numpy, train.py:226-249: numpy + scikit-image dilation rings on the host)
is the caller's: pass it as ``weight``; None means unit weights, for which the weighted figures equal the plain ones.
"""
import torch


def luminance(target):
    """train.py:437-438."""
    if target.shape[1] == 3:
        return 0.2989 * target[:, 0:1] + 0.5870 * target[:, 1:2] + 0.1140 * target[:, 2:3]
    return target


def metrics_from_sums(sums):
    """sums = [sum se, sum se*w, sum w, sum charbonnier*w, count] (float64) -> dict with the three per-batch figures the
    reference accumulates (train.py:442, :445, :447; charbonnier_loss :187-192, clamp(min=1e-8) on the weight sum)."""
    s = [float(v) for v in sums.tolist()]
    return {"mse": s[0] / max(1.0, s[4]), "mse_weighted": s[1] / max(1e-8, s[2]), "charbonnier": s[3] / max(1e-8, s[2])}


class ValidationAccumulator:
    """Epoch aggregation of train.py:422-426 and :469-472 (means over batches)."""

    def __init__(self):
        self.mse_sum = 0.0
        self.mse_weighted_sum = 0.0
        self.loss_sum = 0.0
        self.batches = 0

    def update(self, sums):
        m = metrics_from_sums(sums)
        self.mse_sum += m["mse"]
        self.mse_weighted_sum += m["mse_weighted"]
        self.loss_sum += m["charbonnier"]
        self.batches += 1
        return m

    def result(self):
        n = max(1, self.batches)
        return {"val_mse": self.mse_sum / n, "val_mse_weighted": self.mse_weighted_sum / n, "val_loss": self.loss_sum / n,
                "batches": self.batches}


@torch.no_grad()
def validate(model, batches, eps=1e-3):
    """batches: iterable of (target, input[, weight]) CUDA tensors in the loader's order (train.py:434-436).
    Returns the epoch means and the list of per-batch probabilities (for the caller's AUROC/AUPRC, :449-462)."""
    acc = ValidationAccumulator()
    probs = []
    for item in batches:
        target, inp = item[0], item[1]
        weight = item[2] if len(item) > 2 else None
        _, prob, sums = model.forward_eval(inp, target, weight=weight, eps=eps)
        acc.update(sums)
        probs.append(prob)
    return acc.result(), probs
