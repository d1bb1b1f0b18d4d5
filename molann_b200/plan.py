"""Compile Feature lists / atom groups into the integer *feature program* the kernels execute.

Index conventions restated from the reference: ``Feature.get_atom_indices()`` is 1-based
(molann/feature.py:123), ``FeatureMap`` subtracts 1 (molann/ann.py:258) and maps global -> local slots
with ``list.index`` (molann/ann.py:261); ``AlignmentLayer`` uses 0-based ``ix`` directly (:131,144).
Program entry = 6 int32 ``{type, a0, a1, a2, a3, out_col}`` (include/molann_b200.h); a position
feature over m atoms becomes m entries, columns atom-major / xyz-minor (molann/ann.py:354).
"""
from typing import List, Sequence, Tuple

import numpy as np

ENTRY_INTS = 6
ANGLE, BOND, DIHEDRAL, POSITION = 0, 1, 2, 3
ACT_IDS = {"Tanh": 0, "ReLU": 1, "Sigmoid": 2, "Identity": 3}


def local_indices(global_idx: Sequence[int], input_atom_indices: List[int], what: str) -> List[int]:
    """Global (0-based) -> local slot inside the input atom group; ValueError like the reference."""
    try:
        return [input_atom_indices.index(int(idx)) for idx in global_idx]
    except ValueError:
        raise ValueError(what)


def feature_dim(type_id: int, n_atoms: int, use_angle_value: bool) -> int:
    """FeatureMap.dim(), molann/ann.py:276-286."""
    if type_id == ANGLE or type_id == BOND:
        return 1
    if type_id == DIHEDRAL:
        return 1 if use_angle_value else 2
    return 3 * n_atoms


def compile_entries(type_id: int, local_idx: Sequence[int], use_angle_value: bool,
                    col0: int = 0) -> Tuple[np.ndarray, int]:
    """Program entries of ONE feature starting at output column ``col0`` -> (entries[E,6], dim)."""
    dim = feature_dim(type_id, len(local_idx), use_angle_value)
    if type_id == POSITION:
        ent = np.zeros((len(local_idx), ENTRY_INTS), dtype=np.int32)
        for j, a in enumerate(local_idx):
            ent[j] = (POSITION, a, 0, 0, 0, col0 + 3 * j)
        return ent, dim
    atoms = list(local_idx) + [0] * (4 - len(local_idx))
    ent = np.asarray([[type_id, atoms[0], atoms[1], atoms[2], atoms[3], col0]], dtype=np.int32)
    return ent, dim


def compile_feature_program(features: Sequence[Tuple[int, Sequence[int]]],
                            use_angle_value: bool) -> Tuple[np.ndarray, int]:
    """Concatenate feature programs in list order (column order of molann/ann.py:473)."""
    chunks, col = [], 0
    for type_id, local_idx in features:
        ent, dim = compile_entries(type_id, local_idx, use_angle_value, col)
        chunks.append(ent)
        col += dim
    return np.concatenate(chunks, axis=0).astype(np.int32), col
