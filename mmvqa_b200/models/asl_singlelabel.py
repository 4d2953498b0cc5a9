"""ASLSingleLabel loss.  Mirrors models/asl_singlelabel.py:9-52 of the reference."""
import torch.nn as nn

from .. import functional as Fn


class ASLSingleLabel(nn.Module):
    """Asymmetric single-label loss.  One kernel computes log-softmax, the asymmetric focusing weights,
    label smoothing, the per-sample loss and d(loss)/d(logits)."""

    def __init__(self, gamma_pos=0, gamma_neg=4, eps: float = 0.1, reduction='mean'):
        super(ASLSingleLabel, self).__init__()
        self.eps = eps
        self.targets_classes = []
        self.gamma_pos = gamma_pos
        self.gamma_neg = gamma_neg
        self.reduction = reduction

    def forward(self, inputs, target):
        """inputs: (batch_size, number_classes) float32/bfloat16 CUDA; target: (batch_size,)"""
        if inputs.dim() != 2:
            raise ValueError("ASLSingleLabel expects (batch_size, number_classes) logits")
        loss_rows, tc = Fn.ASLFn.apply(inputs, target, float(self.gamma_pos), float(self.gamma_neg), float(self.eps), True)
        self.targets_classes = tc          # smoothed one-hot, as the reference leaves it
        if self.reduction == 'mean':
            return loss_rows.mean()
        return loss_rows
