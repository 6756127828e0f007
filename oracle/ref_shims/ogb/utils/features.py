"""Feature vocabulary sizes of ogb.utils.features (ogb 1.3.1): allowable_features lengths."""


def get_atom_feature_dims():
    # atomic_num(118+misc), chirality, degree(0..10+misc), formal_charge(-5..5+misc), numH(0..8+misc),
    # number_radical_e(0..4+misc), hybridization(5+misc), is_aromatic, is_in_ring
    return [119, 4, 12, 12, 10, 6, 6, 2, 2]


def get_bond_feature_dims():
    # bond_type(4+misc), bond_stereo(6), is_conjugated
    return [5, 6, 2]
