"""Utterance sharding for multi-GPU inference: utterances are independent end to end, so a
batch is split into contiguous, balanced shards, one per rank, with NO collective on the data path.  Rank r
decodes utterances [lo, hi) and passes `utt_offset = lo`, so the Philox dropout masks (keyed by global
utterance id) are those of the unsharded batch and the sharded result is bit-identical to it."""
from __future__ import annotations

from typing import List, Tuple


def shard_range(n_utts: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced split: the first n_utts % world_size ranks get one extra utterance."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(n_utts, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_ranges(n_utts: int, world_size: int) -> List[Tuple[int, int]]:
    return [shard_range(n_utts, world_size, r) for r in range(world_size)]


def sharded_inference(model, phonemes, phoneme_lens, max_len: int, seed: int, rank: int, world_size: int):
    """Run `model.inference` (either backend: same API) on this rank's shard.  Returns (lo, hi, outputs)."""
    lo, hi = shard_range(phonemes.shape[0], world_size, rank)
    if hi == lo:
        return lo, hi, None
    kw = {"utt_offset": lo} if hasattr(model, "_device_index") else {"utt_ids": list(range(lo, hi))}
    return lo, hi, model.inference(phonemes[lo:hi], phoneme_lens[lo:hi], max_len=max_len, seed=seed, **kw)
