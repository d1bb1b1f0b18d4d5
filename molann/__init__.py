"""Drop-in ``molann`` namespace: ``molann.ann`` and ``molann.feature`` backed by molann_b200."""
