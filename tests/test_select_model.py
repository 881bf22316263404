"""The argument behind the scan's last-CTA selection (pcv_topk.cuh `block_select_lists`), checked on a numpy
model of its steps: given n_lists sorted, 0-padded lists of k distinct keys, the k-th largest list HEAD h is a
lower bound of the k-th largest key overall, so the keys >= h always contain the top-k; when more than
sel_cap keys pass (the best rows all sit in a few lists) the kernel falls back to the radix select.  The kernel
itself is held to the oracle bit for bit by the GPU tests; this pins the reasoning, lists of every shape."""
import numpy as np
import pytest


def select_lists_model(lists: np.ndarray, k: int, sel_cap: int):
    """Returns (top-k keys descending 0-padded, took_fallback).  Mirrors the kernel step by step."""
    n_lists = lists.shape[0]
    allk = lists.reshape(-1)
    if n_lists < 4 * k:
        fallback = True
    else:
        heads = lists[:, 0]
        live_heads = np.sort(heads[heads != 0])[::-1]
        h = live_heads[k - 1] if live_heads.size >= k else 0  # fewer than k non-empty lists: everything survives
        survivors = allk[(allk != 0) & (allk >= h)]
        fallback = survivors.size > sel_cap
        if not fallback:
            top = np.sort(survivors)[::-1][:k]
            return np.pad(top, (0, k - top.size)), False
    top = np.sort(allk[allk != 0])[::-1][:k]
    return np.pad(top, (0, k - top.size)), fallback


def deal(keys: np.ndarray, n_lists: int, k: int, owner: np.ndarray) -> np.ndarray:
    """Each list keeps the k best of the keys its owner index assigns to it (what a CTA's partial list is)."""
    lists = np.zeros((n_lists, k), dtype=np.uint64)
    for c in range(n_lists):
        mine = np.sort(keys[owner == c])[::-1][:k]
        lists[c, :mine.size] = mine
    return lists


@pytest.mark.parametrize("n_lists,k", [(148, 10), (148, 1), (148, 32), (148, 37), (64, 16), (40, 10), (8, 2)])
@pytest.mark.parametrize("dealing", ["even", "one_list", "staircase", "few_lists", "sparse", "empty"])
def test_head_threshold_selection_equals_the_sorted_top_k(n_lists, k, dealing):
    rng = np.random.default_rng(n_lists * 1000 + k)
    n = 20_000
    keys = np.unique(rng.integers(1, 1 << 40, size=n, dtype=np.uint64))
    rng.shuffle(keys)
    n = keys.size
    if dealing == "even":
        owner = rng.integers(0, n_lists, n)
    elif dealing == "one_list":  # the best rows are contiguous: they all land in one CTA
        owner = rng.integers(0, n_lists, n)
        owner[np.argsort(keys)[::-1][:4 * k]] = 3 % n_lists
    elif dealing == "staircase":  # rows sorted by score (or all equal: ties go by id): list c holds the c-th block of ranks
        owner = np.empty(n, dtype=np.int64)
        owner[np.argsort(keys)[::-1]] = np.minimum(np.arange(n) // (n // n_lists + 1), n_lists - 1)
    elif dealing == "few_lists":  # a tiny corpus: only three CTAs saw a row
        owner = rng.integers(0, 3, n)
    elif dealing == "sparse":  # fewer keys than k overall
        keys = keys[:max(1, k // 2)]
        owner = rng.integers(0, n_lists, keys.size)
    else:
        keys = keys[:0]
        owner = np.zeros(0, dtype=np.int64)
    lists = deal(keys, n_lists, k, owner)
    got, fell_back = select_lists_model(lists, k, sel_cap=4 * k)
    want = np.sort(keys)[::-1][:k]
    want = np.pad(want, (0, k - want.size))
    assert np.array_equal(got, want)
    if n_lists >= 4 * k:
        if dealing == "even":
            assert not fell_back, "evenly dealt rows must take the short path"
        if dealing == "staircase" and k > 4:
            assert fell_back, "k lists of k keys pass the head threshold: k*k > 4k, the radix select takes over"


def test_kth_head_is_a_lower_bound_of_the_kth_key():
    rng = np.random.default_rng(7)
    for _ in range(200):
        n_lists, k = int(rng.integers(4, 160)), int(rng.integers(1, 12))
        keys = np.unique(rng.integers(1, 1 << 30, size=int(rng.integers(1, 3000)), dtype=np.uint64))
        rng.shuffle(keys)
        lists = deal(keys, n_lists, k, rng.integers(0, n_lists, keys.size))
        heads = np.sort(lists[:, 0][lists[:, 0] != 0])[::-1]
        if heads.size >= k:
            kth = np.sort(keys)[::-1][k - 1]
            assert heads[k - 1] <= kth
