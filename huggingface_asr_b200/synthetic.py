"""Synthetic CTC-head outputs of the shapes BASELINE.json names.

No dataset or checkpoint is reachable, so every parity test and benchmark runs on these:
  * "peaky": per utterance a transcript of U = T // 8 labels + eos is aligned to sorted
    frame positions; logits = N(0,1) + 8 at (frame, aligned label or blank) -- what a trained
    CTC head looks like, and what makes the joint decode terminate after about U + 1 steps;
  * "flat":  logits = N(0,1) -- numerical stress (every prefix stays plausible).
Frame counts follow the reference front-end (80-dim fbank at 10 ms, two stride-2 convs:
src/reguler/extractors.py:12-16): 10 s -> 248, 15 s -> 373, 30 s -> 748 frames.
Token ids follow src/trainers/train_tokenizer.py:36-50: bos 0, eos 1, unk 2, pad = blank 3.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

BOS, EOS, UNK, PAD = 0, 1, 2, 3
BLANK = PAD


def frames_for_seconds(sec: float, conv_padding: int = 0) -> int:
    """T for an utterance of `sec` seconds (src/models/extractors.py:133-162 length formula)."""
    f = 1 + (int(16000 * sec) - 400) // 160
    for _ in range(2):
        f = (f + 2 * conv_padding - 3) // 2 + 1
    return f


@dataclass
class DecodeConfig:
    name: str
    B: int
    W: int
    T: int
    V: int = 5000
    ctc_weight: float = 0.3
    kind: str = "peaky"
    ragged: bool = False
    n_utts: int | None = None  # total utterances of the job (>= B), defaults to B


# BASELINE.json "configs", as shapes (SURVEY.md section 8d "Configs as shapes").
CONFIGS = {
    "C1": DecodeConfig("ED_small 16 x 10 s, beam 10", B=16, W=10, T=248),
    "C2": DecodeConfig("ED_base 256 x 15 s, beam 10, 5000-token BPE", B=256, W=10, T=373),
    "C3": DecodeConfig("DeCRED_base full-vocab, batch 64, beam 10", B=64, W=10, T=373),
    "C4": DecodeConfig("E-Branchformer 30 s, beam 20, long-T stress", B=32, W=20, T=748),
    "C5": DecodeConfig("8192 utterances sharded over the GPUs, beam 10", B=256, W=10, T=373, n_utts=8192),
}


def make_lengths(B: int, T: int, ragged: bool, gen: torch.Generator) -> torch.Tensor:
    if not ragged:
        return torch.full((B,), T, dtype=torch.long)
    lo = max(1, int(0.6 * T))
    lens = torch.randint(lo, T + 1, (B,), generator=gen)
    lens[0] = T  # at least one full-length utterance, as in a padded batch
    return lens


def _no_repeats_on_adjacent_frames(labels, pos, V):
    """CTC merges a label repeated on two neighbouring frames into one: the aligned transcript would then not be what any
    decoder can recover (1 utterance in ~800 at these sizes).  Such a repeat is replaced by the next token id -- without
    touching the random stream, so every other utterance keeps its data."""
    labels = list(labels)
    for i in range(len(labels) - 1):
        if labels[i + 1] == labels[i] and int(pos[i + 1]) == int(pos[i]) + 1 and labels[i + 1] != EOS:
            labels[i + 1] = 5 + (labels[i + 1] - 5 + 1) % (V - 5)
    return labels


def make_encoder_logits(B: int, T: int, V: int, kind: str = "peaky", ragged: bool = False, seed: int = 20240,
                        device: str | torch.device = "cpu", boost: float = 8.0):
    """Returns (logits (B,T,V) fp32, lens (B,) int64, transcripts: list of label lists incl. eos).

    Always generated with a CPU generator (bit-identical on every box) and then moved.
    """
    gen = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, T, V, generator=gen, dtype=torch.float32)
    lens = make_lengths(B, T, ragged, gen)
    transcripts = []
    if kind == "peaky":
        for b in range(B):
            L = int(lens[b])
            U = max(1, L // 8)
            labels = torch.randint(5, V, (U,), generator=gen).tolist() + [EOS]
            n = len(labels)
            if L - 2 >= n:
                pos = torch.randperm(L - 2, generator=gen)[:n].sort().values + 1
            else:  # very short utterance: as many labels as fit
                labels = labels[: max(1, L - 2)]
                n = len(labels)
                pos = torch.arange(1, n + 1)
            labels = _no_repeats_on_adjacent_frames(labels, pos, V)
            target = torch.full((T,), BLANK, dtype=torch.long)
            target[pos] = torch.tensor(labels)
            logits[b, torch.arange(T), target] += boost
            transcripts.append(labels)
    elif kind == "flat":
        transcripts = [[] for _ in range(B)]
    else:
        raise ValueError(f"unknown kind {kind!r}")
    return logits.to(device), lens.to(device), transcripts


def make_encoder_hidden(B: int, T: int, V: int, d: int = 512, kind: str = "peaky", ragged: bool = False, seed: int = 20240,
                        boost: float = 8.0):
    """Encoder hidden states + CTC head for the N4 boundary: returns (hidden (B,T,d), weight (V,d), bias (V), lens, transcripts)
    such that hidden @ weight.T + bias looks like make_encoder_logits' output: N(0,1) noise plus `boost` at the aligned label.
    The head rows have unit norm, the hidden state of a frame is noise + boost * (row of its label)."""
    gen = torch.Generator().manual_seed(seed)
    weight = torch.randn(V, d, generator=gen, dtype=torch.float32)
    weight = weight / weight.norm(dim=1, keepdim=True)
    bias = 0.1 * torch.randn(V, generator=gen, dtype=torch.float32)
    hidden = torch.randn(B, T, d, generator=gen, dtype=torch.float32)
    lens = make_lengths(B, T, ragged, gen)
    transcripts = []
    if kind == "peaky":
        for b in range(B):
            L = int(lens[b])
            U = max(1, L // 8)
            labels = torch.randint(5, V, (U,), generator=gen).tolist() + [EOS]
            n = len(labels)
            if L - 2 >= n:
                pos = torch.randperm(L - 2, generator=gen)[:n].sort().values + 1
            else:
                labels = labels[: max(1, L - 2)]
                n = len(labels)
                pos = torch.arange(1, n + 1)
            labels = _no_repeats_on_adjacent_frames(labels, pos, V)
            target = torch.full((T,), BLANK, dtype=torch.long)
            target[pos] = torch.tensor(labels)
            hidden[b] += boost * weight[target]
            transcripts.append(labels)
    else:
        transcripts = [[] for _ in range(B)]
    return hidden, weight, bias, lens, transcripts


def make_attention_scores(BW: int, V: int, step: int, seed: int = 7, device: str | torch.device = "cpu",
                          scale: float = 2.0) -> torch.Tensor:
    """Scorer-only stand-in for the attention decoder: log_softmax(scale * N(0,1)), (BW,V)."""
    gen = torch.Generator().manual_seed(seed * 100003 + step)
    s = torch.randn(BW, V, generator=gen, dtype=torch.float32) * scale
    return torch.log_softmax(s, dim=-1).to(device)


class SyntheticDecoder:
    """Stand-in for the attention decoder (model code outside the path).

    log_softmax(noise + boost * onehot(target)), where noise cycles through a small pool of 0.5*N(0,1)
    tensors and target is the utterance's transcript token at this output position (eos past its end) --
    i.e. a decoder that mostly agrees with the CTC head, so that beams end with eos after about U + 1
    steps like a trained model's do.  Deterministic given (seed, transcripts); the same object semantics
    on CPU (oracle / reference arm) and on the GPU.
    """

    def __init__(self, transcripts, num_beams: int, vocab: int, max_length: int, seed: int = 7, device="cpu", pool: int = 8,
                 boost: float = 10.0, noise: float = 0.5, per_row_noise: bool = True):
        """per_row_noise=False: every utterance sees the same (num_beams, vocab) noise, so an utterance's decode does not
        depend on which batch, or which row of it, the utterance lands in (the sharded job compares hypotheses across
        different shardings)."""
        B = len(transcripts)
        self.W, self.V = num_beams, vocab
        dev = torch.device(device)
        gen = torch.Generator(device=dev).manual_seed(seed)
        if per_row_noise:
            self.pool = [noise * torch.randn(B * num_beams, vocab, generator=gen, device=dev, dtype=torch.float32) for _ in range(pool)]
        else:
            self.pool = [(noise * torch.randn(num_beams, vocab, generator=gen, device=dev, dtype=torch.float32)).repeat(B, 1) for _ in range(pool)]
        tgt = torch.full((B, max_length + 1), EOS, dtype=torch.long)
        for b, tr in enumerate(transcripts):
            n = min(len(tr), max_length + 1)
            tgt[b, :n] = torch.tensor(tr[:n], dtype=torch.long)
        self.targets = tgt.to(dev).repeat_interleave(num_beams, dim=0)  # (BW, max_length + 1)
        self.boost = boost
        self.max_length = max_length

    def retarget(self, transcripts) -> "SyntheticDecoder":
        """Same noise pool, new transcripts (a job that decodes many batches of the same shape: the C5 sharded decode)."""
        B = len(transcripts)
        if B * self.W != self.pool[0].shape[0]:
            raise ValueError(f"retarget: {B} transcripts for a pool of {self.pool[0].shape[0] // self.W} utterances")
        tgt = torch.full((B, self.max_length + 1), EOS, dtype=torch.long)
        for b, tr in enumerate(transcripts):
            n = min(len(tr), self.max_length + 1)
            tgt[b, :n] = torch.tensor(tr[:n], dtype=torch.long)
        self.targets = tgt.to(self.pool[0].device).repeat_interleave(self.W, dim=0)
        return self

    def __call__(self, input_ids: torch.Tensor, step: int) -> torch.Tensor:
        logits = self.pool[step % len(self.pool)].clone()
        logits.scatter_add_(1, self.targets[:, step: step + 1], torch.full_like(logits[:, :1], self.boost))
        return torch.log_softmax(logits, dim=-1)
