"""``molann.ann`` drop-in (reference molann/ann.py) -- re-exports the B200-native classes."""
from molann_b200.ann import (AlignmentLayer, FeatureLayer, FeatureMap, MolANN, PreprocessingANN,  # noqa: F401
                             create_sequential_nn)
