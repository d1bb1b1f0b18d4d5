"""Minimal AtomGroup / Universe stand-in (MDAnalysis is optional).

The hot path touches an atom group only through ``.ix`` (0-based global indices),
``.positions`` (fp32 ``[m,3]`` in group order), ``len()``, iteration over hashable
atoms and ``+`` (order-preserving concatenation) -- reference call sites:
molann/ann.py:131-135,255-256 and molann/feature.py:84,123,258.  Real MDAnalysis
groups satisfy the same protocol and are accepted everywhere this class is.

``Universe`` parses fixed-column PDB ``ATOM``/``HETATM`` records (columns 31-54 for
xyz) and supports the small ``bynum`` selection subset the reference's own test
script uses (test/test_molann.py:29,39,50); like MDAnalysis, one ``bynum a b c``
selection returns atoms in sorted order (reference note molann/feature.py:62-69).
"""
from __future__ import annotations

import numpy as np


class AtomGroup(object):
    """Ordered list of atoms of a system, identified by 0-based global index."""

    def __init__(self, ix, all_positions):
        self._ix = np.asarray(ix, dtype=np.int64).reshape(-1)
        self._all_positions = np.asarray(all_positions, dtype=np.float32)
        if self._ix.size and (self._ix.min() < 0 or self._ix.max() >= len(self._all_positions)):
            raise IndexError("atom index out of range")

    @property
    def ix(self):
        return self._ix

    @property
    def positions(self):
        return np.ascontiguousarray(self._all_positions[self._ix], dtype=np.float32)

    def __len__(self):
        return int(self._ix.size)

    def __iter__(self):
        # atoms are hashable by their global index (enough for set(atom_group))
        return iter(int(i) for i in self._ix)

    def __add__(self, other):
        return AtomGroup(np.concatenate([self._ix, np.asarray(other.ix)]), self._all_positions)

    def __getitem__(self, item):
        return AtomGroup(np.atleast_1d(self._ix[item]), self._all_positions)

    def __repr__(self):
        return f"<AtomGroup with {len(self)} atoms>"


class Universe(object):
    """A set of atoms with reference positions (from a PDB file or an array)."""

    def __init__(self, source):
        if isinstance(source, (str, bytes)):
            names, resids, elements, pos = _parse_pdb(source)
        else:
            pos = np.asarray(source, dtype=np.float32).reshape(-1, 3)
            names = ["X"] * len(pos)
            resids = [1] * len(pos)
            elements = ["X"] * len(pos)
        self._positions = np.ascontiguousarray(pos, dtype=np.float32)
        self.names = names
        self.resids = np.asarray(resids, dtype=np.int64)
        self.elements = elements
        self.atoms = AtomGroup(np.arange(len(pos)), self._positions)

    def select_ix(self, ix):
        """Order-preserving group from 0-based indices."""
        return AtomGroup(ix, self._positions)

    def select_atoms(self, selector):
        """Subset of the MDAnalysis selection language.

        Supported: ``bynum i j k`` / ``bynum a:b`` (1-based, result sorted), ``resid r``,
        ``all``, ``not name H*``-style heavy atom selection via ``heavy``.
        """
        tok = selector.strip().split()
        if not tok:
            raise ValueError("empty selection")
        key = tok[0].lower()
        n = len(self.atoms)
        if key == "all":
            return self.select_ix(np.arange(n))
        if key == "bynum":
            picked = []
            for t in tok[1:]:
                if t.lower() in ("and", "or"):
                    continue
                if ":" in t or "-" in t:
                    a, b = t.replace("-", ":").split(":")
                    picked.extend(range(int(a), int(b) + 1))
                else:
                    picked.append(int(t))
            ix = np.unique(np.asarray(picked, dtype=np.int64) - 1)
            return self.select_ix(ix)
        if key == "resid":
            wanted = set(int(t) for t in tok[1:])
            return self.select_ix(np.nonzero([r in wanted for r in self.resids])[0])
        if key == "heavy":
            return self.select_ix(np.nonzero([e != "H" for e in self.elements])[0])
        raise NotImplementedError(
            f"selection '{selector}' needs MDAnalysis; the stand-in supports bynum/resid/all/heavy")


def _parse_pdb(path):
    names, resids, elements, pos = [], [], [], []
    with open(path, "r") as fh:
        for line in fh:
            if not (line.startswith("ATOM") or line.startswith("HETATM")):
                continue
            name = line[12:16].strip()
            names.append(name)
            try:
                resids.append(int(line[22:26]))
            except ValueError:
                resids.append(0)
            pos.append([float(line[30:38]), float(line[38:46]), float(line[46:54])])
            el = line[76:78].strip() if len(line) >= 78 else ""
            if not el:
                el = name.lstrip("0123456789")[:1]
            elements.append(el.upper())
    return names, resids, elements, np.asarray(pos, dtype=np.float32)
