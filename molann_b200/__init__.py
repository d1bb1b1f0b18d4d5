"""molann_b200 -- B200-native implementation of molann's per-frame preprocessing + network hot path.

``molann_b200.ann`` / ``molann_b200.feature`` are drop-ins for ``molann.ann`` / ``molann.feature``
(the top-level ``molann`` package of this repo simply re-exports them).
"""
__version__ = "0.1.0"
