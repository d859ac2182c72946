"""LM shallow fusion next to the CTC rescorer (second half): `LMRescorerLogitsProcessor`.

Reference: src/decoding/shallow_fussion.py:5-58 (copy: src/reguler/modeling_decred.py:36-90).  Same class name, same
constructor `(lm_weight, lm_model, device)`, same `__call__(input_ids, scores) -> scores + lm_weight * log_softmax(LM)`.
The reference re-runs the language model on the WHOLE prefix at every decoding step (:45-55; the `past_key_values` lines are
commented out with a TODO): O(L^2) LM work per utterance.  Here the LM's KV cache is kept across steps and only the
newest token is fed:

  * beam search permutes / duplicates hypotheses between two calls, and a `LogitsProcessor` is not told how (HF hands
    `beam_idx` only to `_reorder_cache`).  When the caller forwards it (`set_beam_idx`, done by
    `JointCTCAttentionGenerationMixin._reorder_cache`) the cache rows are reordered with it; otherwise each row's parent
    is recovered by matching its prefix `input_ids[:, :-1]` against the rows of the previous call (hash match, then an
    exact comparison);
  * whenever the prefix of some row cannot be matched exactly (first call, a new `generate()`, a caller that edits
    prefixes) the processor falls back to the reference's full-prefix forward for that call and rebuilds the cache, so
    the result is always that of the reference up to fp32 rounding of the attention sums.

This is host-side Python around a user-supplied `transformers` causal LM (model code, not part of the CUDA path): it runs
wherever the LM runs.
"""
from __future__ import annotations

import torch
from transformers import LogitsProcessor


class LMRescorerLogitsProcessor(LogitsProcessor):
    """Logits processor that adds `lm_weight * log p_LM(token | prefix)` to the next-token scores (shallow fusion)."""

    def __init__(self, lm_weight: float, lm_model, device, use_cache: bool = True):
        super().__init__()
        self.lm_model = lm_model.to(device)  # reference :11
        self.lm_weight = lm_weight
        self.use_cache = use_cache
        self.past_key_values = None
        self._prev_ids = None   # input_ids of the previous call: the prefixes the cache rows belong to
        self._beam_idx = None   # rows of the previous call that this call's rows continue, if the caller told us
        self.full_forwards = 0  # calls that ran the LM on the whole prefix (1 per generate() when the cache is coherent)
        self.cached_forwards = 0

    def reset(self) -> None:
        self.past_key_values, self._prev_ids, self._beam_idx = None, None, None

    def set_beam_idx(self, beam_idx: torch.Tensor) -> None:
        """Optional: the `beam_idx` of HF's `_reorder_cache` (row j of the next call continues row beam_idx[j])."""
        self._beam_idx = beam_idx

    @staticmethod
    def _row_hash(ids: torch.Tensor) -> torch.Tensor:
        """Polynomial hash of every row (int64 arithmetic wraps; collisions are caught by the exact comparison)."""
        L = ids.shape[1]
        mult = torch.full((L,), 1000003, dtype=torch.long, device=ids.device)
        mult[0] = 1
        weights = torch.cumprod(mult, 0)
        return ((ids + 7) * weights).sum(1)

    def _parents(self, input_ids: torch.Tensor):
        """Row indices into the previous call whose prefixes equal input_ids[:, :-1], or None if the cache cannot be reused."""
        prev = self._prev_ids
        if self.past_key_values is None or prev is None or input_ids.shape[1] != prev.shape[1] + 1:
            return None
        prefix = input_ids[:, :-1]
        idx = self._beam_idx
        self._beam_idx = None
        if idx is not None and idx.numel() == input_ids.shape[0]:
            idx = idx.to(device=input_ids.device, dtype=torch.long)
            if int(idx.max()) < prev.shape[0] and bool((prev[idx] == prefix).all()):
                return idx
        hp, hc = self._row_hash(prev), self._row_hash(prefix)
        R, P = hc.shape[0], hp.shape[0]
        if R == P and bool((prev == prefix).all()):  # greedy search / no reordering
            return torch.arange(R, device=input_ids.device)
        idx = torch.empty(R, dtype=torch.long, device=input_ids.device)
        for r0 in range(0, R, 4096):  # bounded (R, P) comparison blocks
            idx[r0:r0 + 4096] = (hc[r0:r0 + 4096, None] == hp[None, :]).to(torch.uint8).argmax(1)
        if not bool((prev[idx] == prefix).all()):
            return None
        return idx

    def __call__(self, input_ids: torch.LongTensor, scores: torch.FloatTensor) -> torch.FloatTensor:
        with torch.no_grad():
            parents = self._parents(input_ids) if self.use_cache else None
            if parents is None:
                outputs = self.lm_model(input_ids, use_cache=self.use_cache)  # reference :45-52
                self.full_forwards += 1
            else:
                cache = self.past_key_values
                if not torch.equal(parents, torch.arange(parents.numel(), device=parents.device)):
                    cache.reorder_cache(parents)
                outputs = self.lm_model(input_ids[:, -1:], past_key_values=cache, use_cache=True)
                self.cached_forwards += 1
            if self.use_cache:
                self.past_key_values = outputs.past_key_values
                self._prev_ids = input_ids.clone()
            lm_scores = torch.nn.functional.log_softmax(outputs.logits[:, -1, :], dim=-1)  # :54
        return scores + self.lm_weight * lm_scores  # :55
