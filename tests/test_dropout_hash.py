"""CPU: the counter-based dropout specification shared by the attention and streaming kernels
(csrc/common.cuh: drop_row_key / drop_hash_pair / dropout_threshold).  The vectorised numpy restatements the GPU
tests use (atq/attention.py) are pinned here against a scalar, integer-only statement of the same hash, so the
three descriptions (CUDA, numpy, this file) cannot drift apart silently."""
import numpy as np
import pytest

from atq import attention as A

M32 = 0xFFFFFFFF


def _row_key(seed, row_id):
    x = ((row_id * 0x9E3779B1) & M32) ^ (seed & M32)
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & M32
    x ^= x >> 13
    return (x + ((seed >> 32) & M32)) & M32


def _pair_hash(key, pair):
    x = (key + pair * 0xC2B2AE35) & M32
    x ^= x >> 16
    x = (x * 0x7FEB352D) & M32
    x ^= x >> 15
    x = (x * 0x846CA68B) & M32
    x ^= x >> 16
    return x


def _keep(key, elem, thresh):
    h = _pair_hash(key, elem // 2)
    return ((h >> 16) if elem % 2 else (h & 0xFFFF)) >= thresh


@pytest.mark.parametrize("p", [0.1, 0.25, 0.5])
def test_attention_mask_matches_scalar_statement(p):
    seed, b, h, l = (7 << 32) | 12345, 2, 3, 37
    keep, p_eff = A.dropout_keep_mask(seed, b, h, l, p)
    thresh = min(max(int(p * 65536.0 + 0.5), 1), 65535)
    assert p_eff == thresh / 65536.0 and abs(p_eff - p) < 1e-4
    for (bi, hi, q, k) in [(0, 0, 0, 0), (0, 0, 0, 1), (1, 2, 36, 36), (1, 0, 5, 20), (0, 1, 17, 3)]:
        row_id = (bi * h + hi) * l + q
        assert keep[bi, hi, q, k] == _keep(_row_key(seed, row_id), k, thresh)
    assert abs(keep.mean() - (1 - p_eff)) < 0.03


@pytest.mark.parametrize("stream_id", [0x0FF1CE, 0x6A7ED])
def test_flat_mask_matches_scalar_statement(stream_id):
    seed, n, p = 987654321987, 1001, 0.2
    keep, p_eff = A.dropout_keep_mask_flat(seed, n, p, stream_id=stream_id)
    thresh = min(max(int(p * 65536.0 + 0.5), 1), 65535)
    key = _row_key(seed, stream_id)
    for e in (0, 1, 2, 3, 500, 999, 1000):
        assert keep[e] == _keep(key, e, thresh)
    assert keep.shape == (n,) and abs(keep.mean() - (1 - p_eff)) < 0.05


def test_no_dropout_keeps_everything():
    keep, p_eff = A.dropout_keep_mask(1, 1, 1, 9, 0.0)
    assert keep.all() and p_eff == 0.0
    keep, p_eff = A.dropout_keep_mask_flat(1, 10, 0.0)
    assert keep.all() and p_eff == 0.0
