"""Feature description of a molecular system (host-side metadata only).

Drop-in for ``molann.feature`` (reference molann/feature.py:25-290): a ``Feature`` is a typed,
ordered atom-index record; the CUDA path never sees these objects -- ``molann_b200.plan`` compiles
them into the integer feature program.  Error behaviour mirrors the reference line by line.
"""
import pandas as pd

_TYPE_IDS = {'angle': 0, 'bond': 1, 'dihedral': 2, 'position': 3}   # reference feature.py:89-97
_ATOM_COUNTS = {'angle': (3, '3 atoms are needed to define an angle feature, {} provided'),
                'bond': (2, '2 atoms are needed to define a bond length feature, {} provided'),
                'dihedral': (4, '4 atoms are needed to define a dihedral angle feature, {} provided')}


class Feature(object):
    r"""Feature of system (reference feature.py:79-137).

    :param str name: feature's name
    :param str feature_type: 'angle', 'bond', 'dihedral' or 'position'
    :param atom_group: MDAnalysis ``AtomGroup`` or :class:`molann_b200.atomgroup.AtomGroup`

    :raises NotImplementedError: unknown feature type
    :raises IndexError: repeated atoms in the group
    :raises AssertionError: atom count does not match the type (3 / 2 / 4)
    """

    def __init__(self, name, feature_type, atom_group):
        if feature_type not in _TYPE_IDS:
            raise NotImplementedError(f'feature {feature_type} not implemented!')
        if len(set(atom_group)) < len(atom_group):
            raise IndexError('atom group contains repeated elements!')
        if feature_type in _ATOM_COUNTS:
            need, msg = _ATOM_COUNTS[feature_type]
            assert len(atom_group) == need, msg.format(len(atom_group))
        self.name = name
        self.type_name = feature_type
        self.atom_group = atom_group
        self.type_id = _TYPE_IDS[feature_type]

    def get_name(self):
        return self.name

    def get_type(self):
        return self.type_name

    def get_atom_indices(self):
        """numpy array of int, (1-based) indices of atoms in the atom group (reference feature.py:123)."""
        return self.atom_group.ix + 1

    def get_type_id(self):
        return self.type_id

    def get_feature_info(self):
        """One-row :class:`pandas.DataFrame` (columns as reference feature.py:137)."""
        return pd.DataFrame({'name': self.name, 'type': self.type_name, 'type_id': self.type_id,
                             'atom indices (1-based)': [self.get_atom_indices()]})


class FeatureFileReader(object):
    r"""Read a list of :class:`Feature` from one ``[section] ... [End]`` block of a text file.

    File format as reference feature.py:147-161: comma separated ``name, type, selector[, selector..]``
    (selectors concatenated in order), ``#`` comment lines.  ``universe`` only needs
    ``select_atoms(str)`` returning groups that support ``+``.
    """

    def __init__(self, feature_file, section_name, universe):
        self.feature_file = feature_file
        self.section_name = section_name
        self.u = universe
        self.feature_list = []

    def read(self):
        self.feature_list = []
        in_section = False
        with open(self.feature_file, "r") as fh:
            for raw in fh:
                line = raw.strip()
                if not line or line.startswith("#"):
                    continue
                if line.startswith("["):
                    tag = line.strip('[]')
                    if tag == self.section_name:
                        in_section = True
                        continue
                    if in_section and tag == 'End':
                        break
                if not in_section:
                    continue
                # note (reference quirk, feature.py:244-253): a stray "[X]" line inside the active
                # section is parsed as a feature line, exactly like the reference does.
                feature_name, feature_type, *selector_list = line.split(',')
                ag = None
                for selector in selector_list:
                    picked = self.u.select_atoms(selector)
                    ag = picked if ag is None else ag + picked
                self.feature_list.append(Feature(feature_name.strip(), feature_type.strip(), ag))
        return self.feature_list

    def get_feature_list(self):
        return self.feature_list

    def get_num_of_features(self):
        return len(self.feature_list)

    def get_feature_info(self):
        df = pd.DataFrame()
        for f in self.feature_list:
            df = pd.concat([df, f.get_feature_info()], ignore_index=True)
        return df
