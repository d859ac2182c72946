"""Utterance sharding for multi-GPU decoding.

Utterances are independent units: the scorer state, posteriors and beams of utterance b never touch those
of another one (every index in the reference is per-b: ctc_scorer.py:55-56,96,191,198).  So the job is split
by utterance, one process per GPU, with NO collective while decoding; the only communication is one
all_gather of the finished hypotheses at the end (NCCL on GPUs, gloo in the CPU tests).

The reference itself shards evaluation the same way, through HF Trainer's distributed sampler
(src/utilities/general_utils.py:151 -> Seq2SeqTrainer.predict) and gathers with `_nested_gather`
(src/utilities/training_utils.py:357-361).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


def shard_utterances(lengths: Sequence[int], world_size: int) -> list[list[int]]:
    """Deal utterances, longest first, round-robin over ranks: balances frames (= scorer work) per rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[r::world_size] for r in range(world_size)]


def make_batches(indices: Sequence[int], batch_size: int) -> list[list[int]]:
    """Consecutive runs of a length-sorted shard: each batch pads to its own longest utterance."""
    return [list(indices[i:i + batch_size]) for i in range(0, len(indices), batch_size)]


def decode_shard(indices: Sequence[int], batch_size: int, load_batch: Callable, decode_batch: Callable, max_length: int,
                 pad: int, device) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Decode this rank's utterances batch by batch.

    load_batch(list of utterance ids) -> whatever decode_batch needs (e.g. (logits, lens, decoder));
    decode_batch(...) -> BeamSearchOutput.  Returns (ids, sequences (n,max_length), lengths, scores) on `device`.
    """
    seqs, lens, scores = [], [], []
    for batch in make_batches(indices, batch_size):
        out = decode_batch(*load_batch(batch))
        s = out.sequences
        if s.shape[1] < max_length:
            s = torch.nn.functional.pad(s, (0, max_length - s.shape[1]), value=pad)
        seqs.append(s[:, :max_length])
        lens.append(out.lengths)
        scores.append(out.scores)
    ids = torch.tensor(list(indices), dtype=torch.long, device=device)
    if not seqs:
        return (ids, torch.empty((0, max_length), dtype=torch.long, device=device), torch.empty((0,), dtype=torch.long, device=device),
                torch.empty((0,), dtype=torch.float32, device=device))
    return ids, torch.cat(seqs), torch.cat(lens), torch.cat(scores)


def gather_hypotheses(ids: torch.Tensor, seqs: torch.Tensor, lens: torch.Tensor, scores: torch.Tensor, n_total: int, pad: int,
                      group=None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The one collective of the path: all_gather of (ids, sequences, lengths, scores), padded to the largest shard,
    scattered back into utterance order.  Every rank returns the full (n_total, max_length) result."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        order = torch.argsort(ids)
        return seqs[order], lens[order], scores[order]
    world = dist.get_world_size(group)
    dev = seqs.device
    n_local = torch.tensor([ids.numel()], dtype=torch.long, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    n_max = int(max(int(c) for c in counts))
    max_length = seqs.shape[1]

    def padded(t, fill):
        out = torch.full((n_max,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=dev)
        out[: t.shape[0]] = t
        return out

    # one int64 payload (ids | lengths | sequences) and one fp32 payload: two collectives in total
    payload = torch.cat([padded(ids, -1).view(n_max, 1), padded(lens, 0).view(n_max, 1), padded(seqs, pad)], dim=1).contiguous()
    all_payload = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(all_payload, payload, group=group)
    sc = padded(scores, 0.0).contiguous()
    all_sc = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(all_sc, sc, group=group)

    out_seqs = torch.full((n_total, max_length), pad, dtype=torch.long, device=dev)
    out_lens = torch.zeros((n_total,), dtype=torch.long, device=dev)
    out_scores = torch.zeros((n_total,), dtype=torch.float32, device=dev)
    for r in range(world):
        n = int(counts[r])
        if n == 0:
            continue
        p = all_payload[r][:n]
        idx = p[:, 0]
        out_seqs[idx] = p[:, 2:]
        out_lens[idx] = p[:, 1]
        out_scores[idx] = all_sc[r][:n]
    return out_seqs, out_lens, out_scores
