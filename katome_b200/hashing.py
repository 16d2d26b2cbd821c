"""Host-side (numpy) mirror of the device key functions in csrc/common.cuh: canonical form,
the two 32-bit placement hashes and the owner rank of a k-mer.  Used by the host routing logic and its
CPU tests; the insert path itself only exists on the GPU."""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
_SALT = np.uint64(0x6B61746F6D65)


def fmix64(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    return z


def _rev2(x: np.ndarray) -> np.ndarray:
    """reverse the 32 two-bit symbols of every u64"""
    x = np.asarray(x, dtype=np.uint64)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF),
                  (16, 0x0000FFFF0000FFFF)):
        x = ((x >> np.uint64(sh)) & np.uint64(m)) | ((x & np.uint64(m)) << np.uint64(sh))
    return (x >> np.uint64(32)) | (x << np.uint64(32))


def revcomp(hi: np.ndarray, lo: np.ndarray, k: int):
    """reverse complement of k-mers given as (hi, lo) u64 pairs (hi = 0 for k <= 32)"""
    hi = np.asarray(hi, dtype=np.uint64)
    lo = np.asarray(lo, dtype=np.uint64)
    if k <= 32:
        return np.zeros_like(lo), _rev2(~lo) >> np.uint64(64 - 2 * k)
    rhi, rlo = _rev2(~lo), _rev2(~hi)  # 128-bit value (rhi:rlo), to be shifted right by 128-2k
    s = 128 - 2 * k
    if s == 0:
        return rhi, rlo
    out_lo = (rlo >> np.uint64(s)) | (rhi << np.uint64(64 - s))
    return rhi >> np.uint64(s), out_lo


def canonical(hi, lo, k: int):
    hi = np.asarray(hi, dtype=np.uint64)
    lo = np.asarray(lo, dtype=np.uint64)
    rhi, rlo = revcomp(hi, lo, k)
    less = (rhi < hi) | ((rhi == hi) & (rlo < lo))
    return np.where(less, rhi, hi), np.where(less, rlo, lo)


def mix64(x: np.ndarray) -> np.ndarray:
    """common.cuh mix64: two multiply / fold-by-32 rounds (control paths, node owners)"""
    x = np.asarray(x, dtype=np.uint64).copy()
    with np.errstate(over="ignore"):
        x *= np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(32)
        x *= np.uint64(0xD6E8FEB86659FD93)
        x ^= x >> np.uint64(32)
    return x


_M32 = np.uint64(0xFFFFFFFF)


def _mix32(x: np.ndarray, c1: int, s2: int, c2: int) -> np.ndarray:
    """32-bit finalizer on u64 lanes holding 32-bit values"""
    x = x & _M32
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(c1)) & _M32
    x = x ^ (x >> np.uint64(s2))
    x = (x * np.uint64(c2)) & _M32
    return x ^ (x >> np.uint64(16))


def _words(hi, lo):
    hi = np.asarray(hi, dtype=np.uint64)
    lo = np.asarray(lo, dtype=np.uint64)
    return lo & _M32, lo >> np.uint64(32), hi & _M32, hi >> np.uint64(32)


def place_hash(hi, lo, k: int) -> np.ndarray:
    """KeyTraits<K>::place_hash of common.cuh: owner rank and sub-table come from it"""
    w0, w1, w2, w3 = _words(hi, lo)
    with np.errstate(over="ignore"):
        if k <= 32:
            f = w1 + w0 * np.uint64(0x85EBCA77)
        else:
            f = w3 + w2 * np.uint64(0x85EBCA77) + w1 * np.uint64(0xC2B2AE3D) + w0 * np.uint64(0x27D4EB2F)
    return _mix32(f, 0x7FEB352D, 15, 0x846CA68B)


def slot_hash(hi, lo, k: int) -> np.ndarray:
    """KeyTraits<K>::slot_hash of common.cuh: home slot inside the sub-table (its high bits: the page)"""
    w0, w1, w2, w3 = _words(hi, lo)
    with np.errstate(over="ignore"):
        f = w0 + w1 * np.uint64(0x9E3779B1)
        if k > 32:
            f = f + w2 * np.uint64(0x85EBCA77) + w3 * np.uint64(0xC2B2AE3D)
    return _mix32(f, 0x85EBCA6B, 13, 0xC2B2AE35)


def owner_of(hi, lo, k: int, world: int, reverse_complement: bool = True) -> np.ndarray:
    """owner rank = place_hash range-reduced to [0, world) (place_of in common.cuh)"""
    if reverse_complement:
        hi, lo = canonical(hi, lo, k)
    h = place_hash(hi, lo, k)
    return ((h * np.uint64(world)) >> np.uint64(32)).astype(np.int64)
