"""SupCon / SimCLR loss.  Mirrors models/SupConLoss/loss.py:11-98 of the reference."""
import torch
import torch.nn as nn

from ... import functional as Fn
from ...config import compute_dtype


class SupConLoss(nn.Module):
    """Supervised contrastive loss (also SimCLR when labels and mask are None).

    The anchor x contrast similarity matrix is one tensor-core GEMM; max / masked exp-sum /
    positive mean / gradient are one fused row kernel.  ``gathered`` (set by
    mmvqa_b200.parallel.gather_supcon_features) is handled outside: this module always sees the
    full contrast set it is given."""

    def __init__(self, temperature=0.07, contrast_mode='all', base_temperature=0.07):
        super(SupConLoss, self).__init__()
        self.temperature = temperature
        self.contrast_mode = contrast_mode
        self.base_temperature = base_temperature

    def forward(self, features, labels=None, mask=None):
        if len(features.shape) < 3:
            raise ValueError('`features` needs to be [bsz, n_views, ...],'
                             'at least 3 dimensions are required')
        if len(features.shape) > 3:
            features = features.view(features.shape[0], features.shape[1], -1)
        batch_size = features.shape[0]
        if labels is not None and mask is not None:
            raise ValueError('Cannot define both `labels` and `mask`')
        elif labels is None and mask is None:
            mask = None                                  # identity mask is generated in the kernel
        elif labels is not None:
            labels = labels.contiguous().view(-1, 1)
            if labels.shape[0] != batch_size:
                raise ValueError('Num of labels does not match num of features')
            mask = torch.eq(labels, labels.T).float().to(features.device)
        else:
            mask = mask.float().to(features.device)
        contrast_count = features.shape[1]
        # view-major concat (loss.py:58): all view-0 rows, then all view-1 rows
        contrast_feature = features.transpose(0, 1).reshape(contrast_count * batch_size, -1)
        if self.contrast_mode == 'one':
            anchor_count = 1
        elif self.contrast_mode == 'all':
            anchor_count = contrast_count
        else:
            raise ValueError('Unknown mode: {}'.format(self.contrast_mode))
        if contrast_feature.dtype not in (torch.float32, torch.bfloat16):
            contrast_feature = contrast_feature.float()
        loss_rows = Fn.SupConFn.apply(contrast_feature.contiguous(), mask, batch_size, anchor_count * batch_size,
                                      float(self.temperature), float(self.base_temperature), compute_dtype())
        return loss_rows.view(anchor_count, batch_size).mean()
