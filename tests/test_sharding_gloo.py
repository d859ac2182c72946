"""world_size-2 gloo test of the multi-GPU host logic: shard by utterance, decode independently (CPU oracle as the
scorer -- allowed in tests), gather the hypotheses with the one collective of the path, compare with a single-rank run."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from huggingface_asr_b200.sharding import decode_shard, gather_hypotheses, make_batches, shard_utterances

N, W, T, V, MAXLEN = 6, 3, 40, 64, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _decode_all(indices, lens_all, seed=77):
    from huggingface_asr_b200.beam_search import joint_beam_search
    from huggingface_asr_b200.synthetic import BLANK, BOS, EOS, SyntheticDecoder, make_encoder_logits
    from oracle import oracle as orc

    logits_all, _, tr_all = make_encoder_logits(N, T, V, "peaky", True, seed=seed)

    def load(batch):
        idx = torch.tensor(batch)
        return logits_all[idx].clone(), lens_all[idx].clone(), [tr_all[i] for i in batch]

    def decode(logits, lens, trs):
        B = logits.shape[0]
        proc = orc.OracleCTCRescorerLogitsProcessor(logits, lens, BLANK, EOS, 0, 0.3, W)
        return joint_beam_search(proc, SyntheticDecoder(trs, W, V, MAXLEN, seed=3, noise=0.0), B, W, V, BOS, EOS, BLANK, max_length=MAXLEN)

    return decode_shard(indices, 2, load, decode, MAXLEN, BLANK, "cpu")


def _worker(rank, world, port, lens_all, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    shards = shard_utterances(lens_all.tolist(), world)
    ids, seqs, lens, scores = _decode_all(shards[rank], lens_all)
    out = gather_hypotheses(ids, seqs, lens, scores, N, 3)
    if rank == 0:
        ret.put(tuple(t.clone() for t in out))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_utterances_balances_and_covers():
    lens = [10, 50, 30, 20, 40, 60, 5]
    shards = shard_utterances(lens, 3)
    assert sorted(i for s in shards for i in s) == list(range(7))
    assert shards[0][0] == 5 and shards[1][0] == 1 and shards[2][0] == 4  # longest first, dealt round-robin
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert make_batches(list(range(5)), 2) == [[0, 1], [2, 3], [4]]
    assert shard_utterances([], 2) == [[], []]


def test_gather_single_process_is_a_reorder():
    ids = torch.tensor([2, 0, 1])
    seqs = torch.arange(12).view(3, 4)
    out = gather_hypotheses(ids, seqs, torch.tensor([4, 4, 4]), torch.tensor([0.2, 0.0, 0.1]), 3, 3)
    assert out[0].tolist() == [[4, 5, 6, 7], [8, 9, 10, 11], [0, 1, 2, 3]]
    assert out[2].tolist() == pytest.approx([0.0, 0.1, 0.2])


@pytest.mark.timeout(300)
def test_two_rank_gloo_decode_matches_single_rank():
    from huggingface_asr_b200.synthetic import make_encoder_logits

    _, lens_all, _ = make_encoder_logits(N, T, V, "peaky", True, seed=77)
    ids, seqs, lens, scores = _decode_all(list(range(N)), lens_all)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lens_all, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (got[0] == seqs).all(), "sharded 1-best sequences differ from the single-rank decode"
    assert (got[1] == lens).all()
    assert (got[2] - scores).abs().max() <= 1e-5
