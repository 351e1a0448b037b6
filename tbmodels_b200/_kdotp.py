"""Host-side mirror of ``tbmodels.kdotp.KdotpModel`` (reference src/tbmodels/kdotp.py:19-100), evaluated on the GPU.

Same constructor contract (every Taylor coefficient must be Hermitian, ``ValueError`` otherwise, :41-45), same
``hamilton(k)`` / ``eigenval(k)`` signatures and return types (:51-100).  ``H(k) = sum_q prod_d k_d^{p_qd} C_q`` has
exactly the shape of the tight-binding build -- real coefficients times Hermitian matrices -- so it runs through the
same DMMA GEMM with a monomial generator in place of the phase generator, then the same batched eigensolver.
"""
from __future__ import annotations

import hashlib
import weakref

import numpy as np

from ._evaluator import Evaluator


def pack_kdotp(taylor_coefficients):
    """``{powers tuple: matrix}`` -> (powers int32 [n_terms, dim], coeff complex128 [n_terms, N, N]) in dict order."""
    keys = list(taylor_coefficients.keys())
    if not keys:
        raise ValueError("a k.p model needs at least one Taylor coefficient")
    dim = len(keys[0])
    powers = np.array(keys, dtype=np.int32).reshape(len(keys), dim)
    coeff = np.stack([np.asarray(taylor_coefficients[k], dtype=np.complex128) for k in keys])
    return np.ascontiguousarray(powers), np.ascontiguousarray(coeff)


class KdotpModel:
    """Duck-type of ``tbmodels.kdotp.KdotpModel``."""

    def __init__(self, taylor_coefficients, device=None) -> None:
        for mat in taylor_coefficients.values():
            if not np.allclose(mat, np.array(mat).T.conj()):
                raise ValueError(f"The provided Taylor coefficient {mat} is not hermitian")
        self.taylor_coefficients = {
            tuple(key): np.array(mat, dtype=complex) for key, mat in taylor_coefficients.items()
        }
        self._device = device
        self._cache = None

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_cache"] = None
        return state

    def evaluator(self) -> Evaluator:
        return kdotp_evaluator_for(self, self)

    def hamilton(self, k):
        return self.evaluator().hamilton(k)

    def eigenval(self, k):
        return self.evaluator().eigenval(k)


def kdotp_evaluator_for(model, holder=None, cache: dict = None, device=None) -> Evaluator:
    """Device evaluator for any object with a ``taylor_coefficients`` dict, rebuilt when the dict content changes."""
    powers, coeff = pack_kdotp(model.taylor_coefficients)
    h = hashlib.blake2b(digest_size=16)
    h.update(powers.tobytes())
    h.update(coeff.tobytes())
    digest = h.digest()
    if holder is not None:
        entry = holder._cache
    else:
        entry = cache.get(id(model))
    if entry is None or entry[0] != digest:
        if entry is not None:
            entry[1].close()
        dev = device if device is not None else getattr(model, "_device", None)
        entry = (digest, Evaluator.from_kdotp(powers, coeff, device=dev))
        if holder is not None:
            holder._cache = entry
        else:
            key = id(model)
            if key not in cache:  # release the device copy when the model object goes away (as evaluator_for does)
                try:
                    weakref.finalize(model, _drop_cached, cache, key)
                except TypeError:  # no weakref support: the entry lives until uninstall()
                    pass
            cache[key] = entry
    return entry[1]


def _drop_cached(cache: dict, key) -> None:
    entry = cache.pop(key, None)
    if entry is not None:
        entry[1].close()
