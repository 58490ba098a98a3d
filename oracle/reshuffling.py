"""ReshufflingBatchSubsampling and SubsampledObjective.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates src/reshuffling.jl:13-60 and
src/algorithms/subsampledobjective.jl:22-90.

`Random.shuffle(rng, dataset)` (reshuffling.jl:29) consumes a Julia RNG stream that
cannot be reproduced here; both the oracle and the product draw the
permutation of shuffle number `k` under key `key` as a Fisher-Yates pass driven by
Philox4x32-10 words with counter (j // 4, k, 0, STREAM_SHUFFLE).  This is the only
integer arithmetic on the path; parity for it is bit-exact.
"""

from __future__ import annotations

import numpy as np

from .philox import philox4x32_10, split_key, STREAM_SHUFFLE


def philox_shuffle(dataset: np.ndarray, key: int, shuffle_index: int) -> np.ndarray:
    """Fisher-Yates, descending i, j = (word_i * (i + 1)) >> 32, word_i = word number i."""
    n = len(dataset)
    nb = (n + 3) // 4
    ctr = np.zeros((nb, 4), dtype=np.uint64)
    ctr[:, 0] = np.arange(nb, dtype=np.uint64)
    ctr[:, 1] = np.uint64(shuffle_index & 0xFFFFFFFF)
    ctr[:, 3] = np.uint64(STREAM_SHUFFLE)
    words = philox4x32_10(ctr, split_key(key)).reshape(-1).astype(np.uint64)
    perm = np.array(dataset, copy=True)
    for i in range(n - 1, 0, -1):
        j = int((words[i] * np.uint64(i + 1)) >> np.uint64(32))
        perm[i], perm[j] = perm[j], perm[i]
    return perm


class ReshufflingBatchSubsampling:
    """reshuffling.jl:13-16."""

    def __init__(self, dataset, batchsize: int):
        self.dataset = np.asarray(dataset)
        self.batchsize = int(batchsize)

    def __len__(self):                                  # :23-25
        return -(-len(self.dataset) // self.batchsize)

    def reshuffle_batches(self, key, shuffle_index):    # :27-32
        shuffled = philox_shuffle(self.dataset, key, shuffle_index)
        b = self.batchsize
        return [(k + 1, shuffled[k * b:(k + 1) * b]) for k in range(len(self))]   # enumerate: 1-based


class ReshufflingState:
    """reshuffling.jl:18-21; `iterator` is the list of remaining (step, batch) pairs and
    `n_shuffles` replaces the position in the RNG stream."""

    def __init__(self, epoch, iterator, n_shuffles, key):
        self.epoch, self.iterator, self.n_shuffles, self.key = epoch, iterator, n_shuffles, key


def sub_init(sub: ReshufflingBatchSubsampling, key: int) -> ReshufflingState:     # :34-36
    return ReshufflingState(1, sub.reshuffle_batches(key, 0), 1, key)


def sub_step(sub: ReshufflingBatchSubsampling, state: ReshufflingState,
             drop_trailing_batch_if_too_small: bool = False):
    """reshuffling.jl:38-60.  Returns (batch, new_state, info)."""
    epoch, iterator, nsh = state.epoch, list(state.iterator), state.n_shuffles
    (sub_step_idx, batch), iterator = iterator[0], iterator[1:]        # Iterators.peel
    if len(iterator) == 0:
        iterator = sub.reshuffle_batches(state.key, nsh)
        nsh += 1
        if drop_trailing_batch_if_too_small and len(batch) < sub.batchsize:
            (sub_step_idx, batch), iterator = iterator[0], iterator[1:]
        epoch += 1
    info = dict(epoch=epoch, step=sub_step_idx)
    return batch, ReshufflingState(epoch, iterator, nsh, state.key), info


# -- src/algorithms/subsampledobjective.jl ---------------------------------------------------
def subsampled_init(sub: ReshufflingBatchSubsampling, key: int) -> ReshufflingState:
    """subsampledobjective.jl:22-45: the state stored is the PRE-step one (:32, :44); the
    probing `step` at :36 advances nothing that is kept (Appendix C.2)."""
    return sub_init(sub, key)


def subsampled_estimate_gradient(sub, state: ReshufflingState, prob, grad_fn):
    """subsampledobjective.jl:64-90.  grad_fn(prob_sub) -> (value, grad, info)."""
    batch, state2, sub_info = sub_step(sub, state, True)               # :79
    prob_sub = prob.subsample(batch)                                   # :80
    value, grad, obj_info = grad_fn(prob_sub)                          # :85
    info = dict(sub_info); info.update(obj_info)                       # :89
    return value, grad, state2, info


def subsampled_estimate_objective(sub, key, prob, obj_fn):
    """subsampledobjective.jl:47-58: average over all length(sub) batches of a fresh epoch
    (short trailing batch included)."""
    state = sub_init(sub, key)
    total = 0.0
    for _ in range(len(sub)):
        batch, state, _ = sub_step(sub, state)
        total += obj_fn(prob.subsample(batch)) / len(sub)
    return total
