"""``molann.feature`` drop-in (reference molann/feature.py)."""
from molann_b200.feature import Feature, FeatureFileReader  # noqa: F401
