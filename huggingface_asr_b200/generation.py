"""The callers of the scorer, for transformers 5.x.

The reference wires its processors into HF `generate()` in `JointCTCAttentionEncoderDecoder`
(src/models/ctc_encoder_plus_autoregressive_decoder.py:360-482; copy: src/reguler/modeling_decred.py:434-482) against
transformers 4.39.3, configures decoding through `GenerationConfigCustom` (src/trainers/train_enc_dec_asr.py:61-85) and
reports decoding speed in `do_evaluate` / `do_generate` (src/utilities/general_utils.py:129-228).  This module restates
those three pieces -- and only those -- for the transformers installed here (5.5), so a model class gets the sm_100a
scorer by inheriting one mixin:

    class MyASRModel(JointCTCAttentionGenerationMixin, SpeechEncoderDecoderModel): ...

  * `JointCTCAttentionGenerationMixin._get_logits_processor`  (reference :360-404, incl. the LM shallow-fusion processor
                                                              :398-404, here with a KV cache: decoding/shallow_fusion.py)
  * `JointCTCAttentionGenerationMixin._reorder_cache`         not in the reference: hands HF's `beam_idx` to the processor
                                                              (used only when the processor was built with use_beam_idx)
  * `joint_ctc_generation_config`                             (GenerationConfigCustom, train_enc_dec_asr.py:61-85)
  * `evaluate_decoding`                                       (do_evaluate's timing / tokens-per-second lines :150-164 and
                                                              the WER that `compute_metrics` gets from jiwer)

Models, trainers, datasets and tokenizers stay out of scope.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, Iterable, Sequence

import torch
from transformers import GenerationConfig, LogitsProcessorList

from .decoding.ctc_scorer import CTCRescorerLogitsProcessor, LogSoftmaxProcessor
from .decoding.shallow_fusion import LMRescorerLogitsProcessor


class GenerationConfigCustom(GenerationConfig):
    """The reference's generation config (src/decoding/config.py:4-23): HF's GenerationConfig plus the joint-decoding
    fields, and `update_from_string` (:25-61) -- the `--override_for_evaluation "ctc_weight=0.3;num_beams=10"` mechanism of
    do_evaluate (src/utilities/general_utils.py:140-147).  `ctc_pre_beam_size` is ours (N2; 0 = full-vocabulary scoring)."""

    def __init__(self, ctc_weight=0.0, ctc_margin=0, lm_weight=0, lm_model=None, space_token_id=-1, eos_space_trick_weight=0,
                 apply_eos_space_trick=False, ctc_pre_beam_size=0, **kwargs):
        super().__init__(**kwargs)
        self.ctc_weight, self.ctc_margin = ctc_weight, ctc_margin
        self.lm_weight, self.lm_model = lm_weight, lm_model
        self.space_token_id = space_token_id
        self.eos_space_trick_weight, self.apply_eos_space_trick = eos_space_trick_weight, apply_eos_space_trick
        self.ctc_pre_beam_size = ctc_pre_beam_size

    _TRUE, _FALSE = ("true", "1", "y", "yes"), ("false", "0", "n", "no")

    def update_from_string(self, update_str: str):
        """`key=value` pairs separated by ';'.  A key must already exist; the new value takes the type the attribute has now
        (bool: true/1/y/yes or false/0/n/no in any case; int; float; str), anything else is refused.  Pairs are applied in
        order: the ones before a failing pair stay applied, as in the reference."""
        pairs = dict(item.split("=") for item in update_str.split(";"))  # a malformed item raises ValueError here, like the reference
        for key, text in pairs.items():
            if not hasattr(self, key):
                raise ValueError(f"key {key} isn't in the original config dict")
            current = getattr(self, key)
            if isinstance(current, bool):  # before int: bool is an int
                low = text.lower()
                if low in self._TRUE:
                    value = True
                elif low in self._FALSE:
                    value = False
                else:
                    raise ValueError(f"can't derive true or false from {text} (key {key})")
            elif isinstance(current, int):
                value = int(text)
            elif isinstance(current, float):
                value = float(text)
            elif isinstance(current, str):
                value = text
            else:
                raise ValueError(f"You can only update int, float, bool or string values in the config, got {text} for key {key}")
            setattr(self, key, value)


def joint_ctc_generation_config(*, ctc_weight: float = 0.0, ctc_margin: int = 0, space_token_id: int = -1,
                                apply_eos_space_trick: bool = False, eos_space_trick_weight: float = 1.0,
                                ctc_pre_beam_size: int = 0, lm_weight: float = 0.0, **generation_kwargs) -> GenerationConfigCustom:
    """GenerationConfigCustom (src/decoding/config.py, train_enc_dec_asr.py:61-85) by keyword, plus `ctc_pre_beam_size` (N2,
    0 = the reference's full-vocabulary scoring).  The reference also stores the language model itself in the config
    (`lm_model`, :73); `generate()` deep-copies its config, so here the LM is handed to the model with `set_lm_model` and
    only `lm_weight` travels in the config (a config that does carry `lm_model` is honoured too)."""
    if ctc_pre_beam_size == 1 or ctc_pre_beam_size < 0 or ctc_pre_beam_size > 64:
        raise ValueError("ctc_pre_beam_size must be 0 (full vocabulary) or in [2, 64]: beam search draws 2W candidates out of W * S")
    return GenerationConfigCustom(ctc_weight=ctc_weight, ctc_margin=ctc_margin, lm_weight=lm_weight, space_token_id=space_token_id,
                                  eos_space_trick_weight=eos_space_trick_weight, apply_eos_space_trick=apply_eos_space_trick,
                                  ctc_pre_beam_size=ctc_pre_beam_size, **generation_kwargs)


def rescale_eval_batch(eval_batch_size: int, beams_before: int, beams_after: int) -> int:
    """do_evaluate's rule (general_utils.py:144-147): when an override changes num_beams, the per-device eval batch shrinks
    (or grows) by the same factor, rounded up, so that batch * beams rows stay within what fitted before."""
    if beams_after == beams_before:
        return eval_batch_size
    return math.ceil(eval_batch_size / (beams_after / beams_before))


def override_for_evaluation(generation_config: GenerationConfigCustom, override: str | None, eval_batch_size: int) -> int:
    """Apply `--override_for_evaluation` (general_utils.py:140-147) to the config in place; returns the eval batch size to use."""
    if not override:
        return eval_batch_size
    before = generation_config.num_beams
    generation_config.update_from_string(override)
    return rescale_eval_batch(eval_batch_size, before, generation_config.num_beams)


class JointCTCAttentionGenerationMixin:
    """Generation hooks of JointCTCAttentionEncoderDecoder for transformers 5.x.

    The model (or its `generate` override, reference :450-482) stores the CTC head's outputs of the current batch in
    `self.encoder_logits` (B,T,V) -- un-expanded, one row per utterance -- and `self.encoder_output_lens` (B,) before
    beam search starts (reference :406-418: `_prepare_encoder_decoder_kwargs_for_generation`); `set_ctc_inputs` does
    that for callers that run the encoder themselves.
    """

    ctc_rescorer_cls = CTCRescorerLogitsProcessor  # tests substitute the CPU oracle's processor here
    log_softmax_cls = LogSoftmaxProcessor
    lm_rescorer_cls = LMRescorerLogitsProcessor
    encoder_logits = None
    encoder_output_lens = None
    ctc_rescorer = None
    lm_rescorer = None
    external_lm = None

    def set_ctc_inputs(self, encoder_logits: torch.Tensor, encoder_output_lens: torch.Tensor) -> None:
        self.encoder_logits, self.encoder_output_lens = encoder_logits, encoder_output_lens

    def set_lm_model(self, lm_model) -> None:
        """The external causal LM for shallow fusion (reference: `generation_config.lm_model`, train_enc_dec_asr.py:73).
        Stored outside the module tree on purpose: it is not a sub-module of the ASR model."""
        object.__setattr__(self, "external_lm", lm_model)

    def _get_logits_processor(self, generation_config, *args, **kwargs) -> LogitsProcessorList:
        processors = super()._get_logits_processor(generation_config, *args, **kwargs)
        if getattr(generation_config, "ctc_weight", 0) and generation_config.ctc_weight > 0:  # reference :382
            if self.encoder_logits is None or self.encoder_output_lens is None:
                raise ValueError("ctc_weight > 0 needs the CTC head outputs: call set_ctc_inputs(encoder_logits, encoder_output_lens) "
                                 "(or store them in _prepare_encoder_decoder_kwargs_for_generation) before generate()")
            if generation_config.num_beams <= 1:  # greedy search hands raw logits to the processors (:383-384)
                processors.append(self.log_softmax_cls())
            extra = {}
            pre_beam = int(getattr(generation_config, "ctc_pre_beam_size", 0) or 0)
            if pre_beam > 0:
                extra["pre_beam_size"] = pre_beam
            eos = generation_config.eos_token_id
            if isinstance(eos, (list, tuple)):
                eos = eos[0]
            elif isinstance(eos, torch.Tensor):
                eos = int(eos.flatten()[0])
            self.ctc_rescorer = self.ctc_rescorer_cls(
                self.encoder_logits,
                self.encoder_output_lens,
                generation_config.pad_token_id,
                eos,
                getattr(generation_config, "ctc_margin", 0),
                generation_config.ctc_weight,
                generation_config.num_beams,
                getattr(generation_config, "space_token_id", -1),
                getattr(generation_config, "apply_eos_space_trick", False),
                getattr(generation_config, "eos_space_trick_weight", 1.0),
                **extra,
            )
            timing = getattr(self, "score_timing", None)  # bench.py: CUDA events around the scorer launches of this generate()
            if timing is not None and hasattr(self.ctc_rescorer, "ctc_prefix_scorer"):
                self.ctc_rescorer.ctc_prefix_scorer._timing = timing
            processors.append(self.ctc_rescorer)
        if getattr(generation_config, "lm_weight", 0) and generation_config.lm_weight > 0:  # reference :398-404
            lm = getattr(generation_config, "lm_model", None) or self.external_lm
            if lm is None:
                raise ValueError("If `lm_weight` is specified, make sure that `lm_model` is defined.")
            self.lm_rescorer = self.lm_rescorer_cls(generation_config.lm_weight, lm, device=self.device)
            processors.append(self.lm_rescorer)
        return processors

    def _reorder_cache(self, past_key_values, beam_idx):
        """HF beam search calls this with the rows the next step continues (global indices b*W + source beam).  The
        reference's processor never sees them (it selects CTC states from token ids only, ctc_scorer.py:326-329); ours
        uses them when it was built with use_beam_idx (always in pre-beam mode)."""
        if self.ctc_rescorer is not None and hasattr(self.ctc_rescorer, "set_beam_idx"):
            self.ctc_rescorer.set_beam_idx(beam_idx)
        if self.lm_rescorer is not None and hasattr(self.lm_rescorer, "set_beam_idx"):
            self.lm_rescorer.set_beam_idx(beam_idx)
        if hasattr(past_key_values, "reorder_cache"):
            past_key_values.reorder_cache(beam_idx)
        return past_key_values

    @torch.no_grad()
    def generate(self, *args, **kwargs):
        try:
            return super().generate(*args, **kwargs)
        finally:  # reference :480-481: nothing of a batch survives its generate()
            self.encoder_logits = None
            self.encoder_output_lens = None
            self.ctc_rescorer = None
            self.lm_rescorer = None


# ------------------------------------------------------------------------------------------------
# evaluation driver
# ------------------------------------------------------------------------------------------------
def edit_distance(ref: Sequence, hyp: Sequence) -> int:
    """Levenshtein distance between two token / word sequences (what jiwer computes for the reference's WER)."""
    prev = list(range(len(hyp) + 1))
    for i, r in enumerate(ref, 1):
        cur = [i] + [0] * len(hyp)
        for j, h in enumerate(hyp, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (r != h))
        prev = cur
    return prev[-1]


def error_rate(refs: Iterable[Sequence], hyps: Iterable[Sequence]) -> float:
    """sum of edit distances / sum of reference lengths (WER over word lists, TER over token lists)."""
    errs = n = 0
    for r, h in zip(refs, hyps):
        errs += edit_distance(r, h)
        n += len(r)
    return errs / max(n, 1)


@dataclass
class DecodingReport:
    """What do_evaluate logs per split (general_utils.py:158-164), plus the error rate when references are given."""
    utterances: int = 0
    tokens_produced: int = 0
    seconds: float = 0.0
    error_rate: float | None = None
    predictions: list = field(default_factory=list)

    @property
    def tokens_per_second(self) -> float:
        return self.tokens_produced / self.seconds if self.seconds > 0 else float("nan")

    @property
    def utterances_per_second(self) -> float:
        return self.utterances / self.seconds if self.seconds > 0 else float("nan")


def evaluate_decoding(model, batches: Iterable[dict], generation_config: GenerationConfig, pad_token_id: int,
                      decode: Callable[[list[int]], Sequence] | None = None,
                      special_token_ids: Sequence[int] = ()) -> DecodingReport:
    """Decode every batch with `model.generate` and report speed and error rate, like do_evaluate / do_generate.

    A batch is a dict of `generate()` keyword arguments; the optional keys "labels" ((B,L) int64, pad- or -100-filled
    references), "encoder_logits" and "encoder_output_lens" are consumed here: the latter two are handed to
    `model.set_ctc_inputs` (a model that computes them itself in `_prepare_encoder_decoder_kwargs_for_generation`, like
    the reference's, just omits them).  `decode` maps a list of token ids to the unit the error rate is counted in
    (e.g. `lambda ids: tokenizer.decode(ids).split()` for WER); default: the token ids themselves.
    Time is wall-clock around generate() with the device synchronised, as in the reference (:150-160)."""
    rep = DecodingReport()
    drop = set(special_token_ids) | {pad_token_id, -100}
    refs, hyps = [], []
    for batch in batches:
        batch = dict(batch)
        labels = batch.pop("labels", None)
        enc_logits, enc_lens = batch.pop("encoder_logits", None), batch.pop("encoder_output_lens", None)
        if enc_logits is not None:
            model.set_ctc_inputs(enc_logits, enc_lens)
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.generate(generation_config=generation_config, **batch)
        seqs = out.sequences if hasattr(out, "sequences") else out
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        rep.seconds += time.perf_counter() - t0
        seqs = seqs.cpu()
        rep.utterances += int(seqs.shape[0])
        rep.tokens_produced += int((seqs != pad_token_id).sum())  # general_utils.py:158
        for row in seqs.tolist():
            ids = [t for t in row if t not in drop]
            rep.predictions.append(ids)
            hyps.append(decode(ids) if decode else ids)
        if labels is not None:
            for row in labels.cpu().tolist():
                ids = [t for t in row if t not in drop]
                refs.append(decode(ids) if decode else ids)
    if refs:
        rep.error_rate = error_rate(refs, hyps)
    return rep


# ------------------------------------------------------------------------------------------------
# n-best generation (do_generate, general_utils.py:186-228) and its dump (save_nbests, generation_utils.py:16-52)
# ------------------------------------------------------------------------------------------------
@dataclass
class NBestOutput:
    """What do_generate collects per split: per batch the (B * group, L) sequences, their (B * group,) scores and the labels."""
    nbests: list = field(default_factory=list)
    scores: list = field(default_factory=list)
    labels: list = field(default_factory=list)
    group_size: int = 1
    eval_batch_size: int | None = None


def generate_nbest(model, batches: Iterable[dict], generation_config: GenerationConfig, num_predictions_to_return: int = 1,
                   eval_beam_factor: int = 1, eval_batch_size: int | None = None) -> NBestOutput:
    """do_generate (general_utils.py:186-228) without the Trainer: return `num_predictions_to_return` hypotheses per utterance
    with their scores.  The config is changed like the reference changes it (:197-201): num_return_sequences,
    return_dict_in_generate, output_scores, num_beams * eval_beam_factor -- and the eval batch size divided by the same factor
    is reported back for whoever builds the batches.  A batch is a dict of generate() kwargs; "labels", "encoder_logits" and
    "encoder_output_lens" are consumed here like in evaluate_decoding."""
    cfg = generation_config
    cfg.num_return_sequences = num_predictions_to_return
    cfg.return_dict_in_generate = True
    cfg.num_beams = cfg.num_beams * eval_beam_factor
    cfg.output_scores = True
    out = NBestOutput(group_size=num_predictions_to_return,
                      eval_batch_size=None if eval_batch_size is None else math.ceil(eval_batch_size / eval_beam_factor))
    for batch in batches:
        batch = dict(batch)
        labels = batch.pop("labels", None)
        enc_logits, enc_lens = batch.pop("encoder_logits", None), batch.pop("encoder_output_lens", None)
        if enc_logits is not None:
            model.set_ctc_inputs(enc_logits, enc_lens)
        res = model.generate(generation_config=cfg, **batch)
        out.nbests.append(res.sequences)
        out.scores.append(res.sequences_scores)
        out.labels.append(labels)
    return out


def save_nbests(path: str, nbests, scores, labels, decode: Callable[[list], str], pad_token_id: int, group_size: int = 1) -> None:
    """The three text files of the reference's save_nbests (generation_utils.py:44-52): `<path>_scores.txt`, `_hyps.txt`,
    `_refs.txt`, one line per hypothesis, `utterance<i>-<rank> <value>` with i counting utterances over all batches and rank
    1..group_size; a reference is repeated for each of its hypotheses; -100 in the labels stands for padding.
    decode(list of ids) -> text (e.g. lambda ids: tokenizer.decode(ids, skip_special_tokens=True))."""
    hyps = [decode(row.tolist()) for batch in nbests for row in batch.unbind()]
    refs = []
    for lab in labels:
        lab = lab.clone()
        lab[lab == -100] = pad_token_id
        refs.extend(decode(row.tolist()) for row in lab.repeat_interleave(group_size, dim=0))
    vals = [float(v) for batch in scores for v in batch.unbind()]
    with open(path + "_scores.txt", "w") as f_scores, open(path + "_hyps.txt", "w") as f_hyps, open(path + "_refs.txt", "w") as f_refs:
        for n, (hyp, val, ref) in enumerate(zip(hyps, vals, refs)):
            name = f"utterance{n // group_size}-{n % group_size + 1}"
            f_scores.write(f"{name} {val}\n")
            f_hyps.write(f"{name} {hyp}\n")
            f_refs.write(f"{name} {ref}\n")
