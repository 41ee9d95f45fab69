"""Shard plan for the tensor-sharded decode mode (BASELINE.json configs[4]) — host logic only.

Not wired into the engine yet (round 1 ships replicas).  The plan is the part that has hard constraints from the quantized
formats, so it is written and tested first:

* column-parallel (split output rows M): wq / wk / wv by heads, w1 / w3 by n_ff; row-parallel (split K): wo and w2, followed
  by one all-reduce each — two exchanges per layer;
* every K-split must fall on a 256-element super-block boundary AND coincide with the producer's row split, so that each
  rank's slice of the activation vector is made of whole Q8_K blocks (the integer dots of a rank are then exactly the
  reference's for those blocks; only the order of the fp32 partial sums across ranks differs);
* heads therefore move in groups of g = 256 / gcd(256, head_dim) ... i.e. whole 256-element blocks of the attention
  output: 2 heads for head_dim 128, 4 for head_dim 64; K/V heads follow their query heads (GQA / MQA: a KV head may be
  replicated on several ranks).
"""
from dataclasses import dataclass
from math import gcd
from typing import List, Tuple

QK_K = 256


def _split_units(n_units: int, parts: int) -> List[int]:
    """n_units as evenly as possible over parts (the first n_units % parts parts get one more)."""
    base, extra = divmod(n_units, parts)
    return [base + (1 if i < extra else 0) for i in range(parts)]


@dataclass
class RankShard:
    rank: int
    heads: Tuple[int, int]          # [first, last) query heads
    kv_heads: Tuple[int, int]       # [first, last) KV heads this rank needs (may overlap other ranks under GQA / MQA)
    attn_k: Tuple[int, int]         # [first, last) elements of the attention output = rows of wq and K-range of wo
    ff: Tuple[int, int]             # [first, last) of n_ff = rows of w1 / w3 and K-range of w2
    vocab: Tuple[int, int]          # [first, last) rows of the output head


def plan(n_embd: int, n_head: int, n_head_kv: int, n_ff: int, n_vocab: int, tp: int) -> List[RankShard]:
    hd = n_embd // n_head
    assert hd * n_head == n_embd and n_head % n_head_kv == 0
    group = QK_K // gcd(QK_K, hd)                 # query heads per 256-element block of the attention output
    if (hd * group) % QK_K or n_head % group:
        raise ValueError("head_dim / n_head do not tile into 256-element blocks")
    head_units = _split_units(n_head // group, tp)
    if n_ff % QK_K:
        raise ValueError("n_ff must be a multiple of 256 for K-quant shards")
    ff_units = _split_units(n_ff // QK_K, tp)
    voc = _split_units(n_vocab, tp)
    per_kv = n_head // n_head_kv
    out, h0, f0, v0 = [], 0, 0, 0
    for r in range(tp):
        h1 = h0 + head_units[r] * group
        f1 = f0 + ff_units[r] * QK_K
        v1 = v0 + voc[r]
        kv0, kv1 = (h0 // per_kv, (h1 - 1) // per_kv + 1) if h1 > h0 else (0, 0)
        out.append(RankShard(r, (h0, h1), (kv0, kv1), (h0 * hd, h1 * hd), (f0, f1), (v0, v1)))
        h0, f0, v0 = h1, f1, v1
    return out


def imbalance(shards: List[RankShard]) -> float:
    """largest rank's mat-vec work relative to the mean (1.0 = perfect)"""
    work = [(s.attn_k[1] - s.attn_k[0]) + (s.ff[1] - s.ff[0]) for s in shards]
    return max(work) / (sum(work) / len(work)) if sum(work) else 1.0
