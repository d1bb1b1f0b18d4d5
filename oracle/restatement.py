"""CPU restatement (PyTorch, fp32 or fp64, autograd-capable) of the molann hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference lines it restates.  Written as pure functions over *local* atom indices so
it can be driven by the same integer feature program the CUDA kernels consume.

Deliberate deviation (documented): the dihedral cross products use
the xyz axis explicitly; the reference's ``torch.cross`` without ``dim`` picks the first
size-3 axis and is wrong for batches of exactly 3 frames.
"""
import torch

ANGLE, BOND, DIHEDRAL, POSITION = 0, 1, 2, 3   # molann/feature.py:87-97


def align_forward(x, local_align_idx, ref_x):
    """Kabsch alignment, molann/ann.py:179-197.  x:[L,n,3] -> [L,n,3]."""
    idx = torch.as_tensor(local_align_idx, dtype=torch.long)
    sel = x[:, idx, :]                                   # :179
    x_c = torch.mean(sel, 1, True)                       # :181
    x_notran = sel - x_c                                 # :183
    prod = torch.matmul(x_notran.permute(0, 2, 1), ref_x.to(x.dtype))   # :185-187
    u, s, vh = torch.linalg.svd(prod)                    # :188
    diag = torch.diag(torch.ones(3, dtype=x.dtype)).unsqueeze(0).repeat(x.size(0), 1, 1)  # :190
    sign_vec = torch.sign(torch.linalg.det(torch.matmul(u, vh))).detach()                # :192
    diag[:, 2, 2] = sign_vec                             # :193
    rot = torch.bmm(torch.bmm(u, diag), vh)              # :195
    return torch.matmul(x - x_c, rot)                    # :197


def _cross(a, b):
    return torch.stack((a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                        a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                        a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]), dim=1)


def feature_forward(x, type_id, idx, use_angle_value):
    """One feature map, molann/ann.py:323-354.  x:[L,n,3] -> [L, dim]."""
    if type_id == ANGLE:                                 # :323-332
        r21 = x[:, idx[0], :] - x[:, idx[1], :]
        r23 = x[:, idx[2], :] - x[:, idx[1], :]
        r21l = torch.norm(r21, dim=1, keepdim=True)
        r23l = torch.norm(r23, dim=1, keepdim=True)
        cos_angle = (r21 * r23).sum(dim=1, keepdim=True) / (r21l * r23l)
        return torch.acos(cos_angle) if use_angle_value else cos_angle
    if type_id == BOND:                                  # :334-336
        r12 = x[:, idx[1], :] - x[:, idx[0], :]
        return torch.norm(r12, dim=1, keepdim=True)
    if type_id == DIHEDRAL:                              # :338-351
        r12 = x[:, idx[1], :] - x[:, idx[0], :]
        r23 = x[:, idx[2], :] - x[:, idx[1], :]
        r34 = x[:, idx[3], :] - x[:, idx[2], :]
        n1 = _cross(r12, r23)
        n2 = _cross(r23, r34)
        cos_phi = (n1 * n2).sum(dim=1, keepdim=True)
        sin_phi = (n1 * r34).sum(dim=1, keepdim=True) * torch.norm(r23, dim=1, keepdim=True)
        radius = torch.sqrt(cos_phi ** 2 + sin_phi ** 2)
        if use_angle_value:
            return torch.atan2(sin_phi, cos_phi)
        return torch.cat((cos_phi / radius, sin_phi / radius), dim=1)
    if type_id == POSITION:                              # :353-354
        ii = torch.as_tensor(list(idx), dtype=torch.long)
        return x[:, ii, :].reshape((-1, len(idx) * 3))
    raise NotImplementedError(type_id)


def feature_layer_forward(x, features, use_angle_value):
    """Concatenate feature maps in list order, molann/ann.py:473.

    ``features`` = list of (type_id, [local atom indices])."""
    return torch.cat([feature_forward(x, t, idx, use_angle_value) for (t, idx) in features], dim=1)


def preprocess_forward(x, local_align_idx, ref_x, features, use_angle_value):
    """PreprocessingANN.forward, molann/ann.py:565 (align_layer may be absent -> Identity, :539-542)."""
    if local_align_idx is not None and len(local_align_idx) > 0:
        x = align_forward(x, local_align_idx, ref_x)
    return feature_layer_forward(x, features, use_angle_value)


def apply_activation(h, act):
    if act == "tanh":
        return torch.tanh(h)
    if act == "relu":
        return torch.relu(h)
    if act == "sigmoid":
        return torch.sigmoid(h)
    if act == "identity":
        return h
    raise NotImplementedError(act)


def mlp_forward(f, weights, biases, act="tanh"):
    """nn.Sequential built by create_sequential_nn, molann/ann.py:62-65 (last layer linear)."""
    h = f
    for k, (w, b) in enumerate(zip(weights, biases)):
        h = torch.nn.functional.linear(h, w.to(h.dtype), b.to(h.dtype))
        if k + 1 < len(weights):
            h = apply_activation(h, act)
    return h


def molann_forward(x, local_align_idx, ref_x, features, use_angle_value, weights, biases, act="tanh"):
    """MolANN.forward, molann/ann.py:624."""
    return mlp_forward(preprocess_forward(x, local_align_idx, ref_x, features, use_angle_value),
                       weights, biases, act)


# ---------------------------------------------------------------------------------------------
# Host-side integer logic (index maps), restated for the bit-exact index tests.
# ---------------------------------------------------------------------------------------------
def local_indices(global_indices_0based, input_atom_indices):
    """molann/ann.py:144 and :261 -- position of each atom inside the input group (list.index)."""
    inp = list(int(i) for i in input_atom_indices)
    return [inp.index(int(g)) for g in global_indices_0based]


def feature_dim(type_id, n_atoms, use_angle_value):
    """molann/ann.py:276-286."""
    if type_id in (ANGLE, BOND):
        return 1
    if type_id == DIHEDRAL:
        return 1 if use_angle_value else 2
    return 3 * n_atoms
