"""Drop-in switch for the reference package: route ``tbmodels.Model.hamilton`` / ``eigenval`` to the GPU.

    import tbmodels, tbmodels_b200
    tbmodels_b200.install()          # Model.hamilton / Model.eigenval now run on the B200
    ...
    tbmodels_b200.uninstall()        # restore the numpy implementations

Signatures, return types and the ``ValueError`` for a bad ``convention`` are the reference's
(src/tbmodels/_tb_model.py:1076-1150), so callers such as the ``tbmodels eigenvals`` CLI
(src/tbmodels/_cli.py:255-257, which passes ``model.eigenval`` as a callable) keep working unchanged.

``Model.hop`` and ``Model.pos`` are public and mutated in place by the reference (``add_hop`` :1215,
``remove_small_hop`` :1254-1256, plain reads of missing ``defaultdict`` keys :206), so every call re-packs
the model (O(n_R N^2), tiny next to the batch) and compares a content digest with the cached device copy.
Device handles live in a side table keyed by ``id(model)`` -- never in ``model.__dict__`` -- so pickling and
HDF5 serialisation of the model are unaffected (reference tests/test_pickle.py, tests/test_hdf5.py).
"""
from __future__ import annotations

import weakref

from ._evaluator import Evaluator
from ._pack import pack_model

_cache: dict = {}  # id(model) -> (digest, Evaluator)
#: how many calls the installed methods have served (evidence for test harnesses that the GPU path really ran)
stats = {"hamilton": 0, "eigenval": 0, "construct_kdotp": 0, "kdotp_hamilton": 0, "kdotp_eigenval": 0}
_originals: dict = {}
_device = None


def _drop(key):
    entry = _cache.pop(key, None)
    if entry is not None:
        entry[1].close()


def evaluator_for(model) -> Evaluator:
    """Cached device evaluator for ``model`` (rebuilt when its hoppings / positions changed)."""
    packed = pack_model(model)
    digest = packed.digest()
    key = id(model)
    entry = _cache.get(key)
    if entry is None or entry[0] != digest:
        if entry is None:
            try:
                weakref.finalize(model, _drop, key)
            except TypeError:  # object without weakref support: the entry lives until uninstall()
                pass
        else:
            entry[1].close()
        entry = (digest, Evaluator(packed, device=_device))
        _cache[key] = entry
    return entry[1]


def _hamilton(self, k, convention=2):
    if convention not in [1, 2]:
        raise ValueError(
            "Invalid value '{}' for 'convention': must be either '1' or '2'".format(convention)
        )
    stats["hamilton"] += 1
    return evaluator_for(self).hamilton(k, convention=convention)


def _eigenval(self, k):
    stats["eigenval"] += 1
    return evaluator_for(self).eigenval(k)


def _construct_kdotp(self, k, order):
    """GPU version of ``Model.construct_kdotp`` (reference :942-982): same result type -- the ``KdotpModel`` class of
    the package the model class comes from (``tbmodels.kdotp.KdotpModel``), else this package's duck-type."""
    import sys

    if order < 0:
        raise ValueError("The order for the k.p model must be positive.")
    stats["construct_kdotp"] += 1
    coeff = evaluator_for(self).construct_kdotp(k, order)
    kdotp_cls = getattr(sys.modules.get(type(self).__module__), "KdotpModel", None)
    if kdotp_cls is None:
        from ._kdotp import KdotpModel as kdotp_cls
    return kdotp_cls(taylor_coefficients=coeff)


def _kdotp_hamilton(self, k):
    from ._kdotp import kdotp_evaluator_for

    stats["kdotp_hamilton"] += 1
    return kdotp_evaluator_for(self, cache=_kdotp_cache, device=_device).hamilton(k)


def _kdotp_eigenval(self, k):
    from ._kdotp import kdotp_evaluator_for

    stats["kdotp_eigenval"] += 1
    return kdotp_evaluator_for(self, cache=_kdotp_cache, device=_device).eigenval(k)


_kdotp_cache: dict = {}


def install_kdotp(kdotp_cls=None, device=None):
    """Replace ``hamilton`` / ``eigenval`` on ``tbmodels.kdotp.KdotpModel`` (reference src/tbmodels/kdotp.py:51-100)."""
    global _device
    if kdotp_cls is None:
        import tbmodels.kdotp

        kdotp_cls = tbmodels.kdotp.KdotpModel
    if device is not None:
        _device = device
    if kdotp_cls not in _originals:
        _originals[kdotp_cls] = (kdotp_cls.__dict__.get("hamilton"), kdotp_cls.__dict__.get("eigenval"))
    kdotp_cls.hamilton = _kdotp_hamilton
    kdotp_cls.eigenval = _kdotp_eigenval
    return kdotp_cls


def install(model_cls=None, device=None):
    """Replace ``hamilton`` / ``eigenval`` on ``tbmodels.Model`` (or on ``model_cls``); with no argument the
    reference's ``KdotpModel`` is switched over as well."""
    global _device
    if model_cls is None:
        import tbmodels  # the reference package; only needed for the drop-in switch

        model_cls = tbmodels.Model
        try:
            install_kdotp(device=device)
        except ImportError:
            pass
    _device = device
    if model_cls not in _originals:
        _originals[model_cls] = (model_cls.__dict__.get("hamilton"), model_cls.__dict__.get("eigenval"),
                                 model_cls.__dict__.get("construct_kdotp"))
    _hamilton.__doc__ = getattr(_originals[model_cls][0], "__doc__", None)
    _eigenval.__doc__ = getattr(_originals[model_cls][1], "__doc__", None)
    model_cls.hamilton = _hamilton
    model_cls.eigenval = _eigenval
    if _originals[model_cls][2] is not None:  # only where the class has the method (reference :942)
        _construct_kdotp.__doc__ = getattr(_originals[model_cls][2], "__doc__", None)
        model_cls.construct_kdotp = _construct_kdotp
    return model_cls


def uninstall(model_cls=None):
    """Restore the original methods and release every cached device copy."""
    classes = [model_cls] if model_cls is not None else list(_originals)
    for cls in classes:
        orig = _originals.pop(cls, None)
        if orig is None:
            continue
        for name, fn in zip(("hamilton", "eigenval", "construct_kdotp"), orig):
            if fn is None:
                if name in cls.__dict__ and (name != "construct_kdotp" or cls.__dict__[name] is _construct_kdotp):
                    delattr(cls, name)
            else:
                setattr(cls, name, fn)
    for key in list(_cache):
        _drop(key)
    for key in list(_kdotp_cache):
        _kdotp_cache.pop(key)[1].close()
