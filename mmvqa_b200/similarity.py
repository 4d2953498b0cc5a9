"""Caption-similarity masks for the SupCon pre-training step.

Mirror of ``SimilarityCalculator.jaccard`` / ``buildMask`` (models/SupConLoss/supcon_utils.py:110-138, 195-199):
the reference fills the [bsz, bsz] mask with two nested Python loops over word sets, O(bsz^2) host work per step
(1 M set intersections at bsz = 1024).  Here the host only turns every document into a sorted array of unique
word ids (O(total words)); the pairwise set arithmetic runs in one kernel (csrc/similarity.cu), bit-exact."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import ops


def encode_word_sets(docs: Sequence[str], vocab: Dict[str, int]) -> List[List[int]]:
    """``set(doc.lower().split())`` (supcon_utils.py:123-124) as sorted unique integer ids; ``vocab`` is shared by
    every document that will be compared and grows as new words appear."""
    rows = []
    for d in docs:
        ids = {vocab.setdefault(w, len(vocab)) for w in d.lower().split()}
        rows.append(sorted(ids))
    return rows


def _pack(rows: List[List[int]], lmax: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    ids = torch.zeros(len(rows), lmax, dtype=torch.int32)
    lens = torch.tensor([len(r) for r in rows], dtype=torch.int32)
    for i, r in enumerate(rows):
        if r:
            ids[i, :len(r)] = torch.tensor(r, dtype=torch.int32)
    return ids.to(device, non_blocking=True), lens.to(device, non_blocking=True)


def jaccard_mask(caption: Sequence[str], aug: Sequence[str], device="cuda") -> torch.Tensor:
    """[len(caption), len(aug)] float32 mask on ``device``: 1 on the diagonal, Jaccard similarity of the two word
    sets elsewhere -- what ``SimilarityCalculator.jaccard(caption, aug, bsz)`` returns for bsz = len(caption)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise ops.L.MMVQAError("jaccard_mask runs on the GPU only: there is no CPU fallback for this path")
    vocab: Dict[str, int] = {}
    ra, rb = encode_word_sets(caption, vocab), encode_word_sets(aug, vocab)
    lmax = max(1, max((len(r) for r in ra + rb), default=1))
    ia, la = _pack(ra, lmax, device)
    ib, lb = _pack(rb, lmax, device)
    return ops.jaccard_mask(ia, la, ib, lb)


def build_mask(bsz: int, caption: Sequence[str], aug: Sequence[str], con_task: str, similarity: str = "jaccard",
               device="cuda"):
    """``buildMask`` (supcon_utils.py:195-199): None for SimCLR, the similarity mask otherwise.  Only the Jaccard
    similarity has a device kernel; the language-model similarities (cosine / sentence_transformers / bert_score)
    are separate models and out of scope."""
    if con_task == "simclr":
        return None
    if similarity not in ("jaccard", "jaccard_similarity"):
        raise NotImplementedError(f"similarity {similarity!r} is computed by an external language model")
    if len(caption) != bsz or len(aug) != bsz:
        raise ValueError("buildMask expects bsz captions and bsz augmentations")
    return jaccard_mask(caption, aug, device)
