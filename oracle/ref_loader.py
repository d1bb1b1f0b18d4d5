"""Locate and import the UNMODIFIED reference (zwpku/molann) if it is available.

Search order: ``/root/reference`` (build container only), ``baseline/_ref`` (the
offline ``pip install --target`` copy that travels to the GPU box).  The reference is
imported under the private module names ``_molann_reference.{ann,feature}`` so it never
collides with this repo's drop-in ``molann`` package.  Returns ``None`` when absent.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = ["/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]
_cache = {}


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """-> (ann_module, feature_module, root_path) or None."""
    if "ref" in _cache:
        return _cache["ref"]
    out = None
    for root in _CANDIDATES:
        ann = os.path.join(root, "molann", "ann.py")
        feat = os.path.join(root, "molann", "feature.py")
        if os.path.isfile(ann) and os.path.isfile(feat):
            try:
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    out = (_load("_molann_reference.ann", ann),
                           _load("_molann_reference.feature", feat), root)
                break
            except Exception:  # pragma: no cover - broken copy
                out = None
    _cache["ref"] = out
    return out
