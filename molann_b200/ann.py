r"""Artificial neural networks for molecular systems -- B200-native drop-in for ``molann.ann``.

Same classes, constructors, attributes, ``state_dict`` keys and TorchScript behaviour as the reference
(zwpku/molann, molann/ann.py); every ``forward`` is ONE call into a dispatcher-registered custom op
(``torch.ops.molann_b200.*``, molann_b200/csrc/torch_shim.cpp) that runs hand-written sm_100a kernels
through the C ABI of include/molann_b200.h.  There is no CPU implementation: inputs must be contiguous
float32 CUDA tensors, anything else raises.

At construction each layer compiles its atom indices into int32 *program* buffers (registered
``persistent=False`` so ``state_dict()`` stays identical to the reference's, and moved by ``.to(device)``).
"""
from typing import List, Optional

import pandas as pd
import torch

from . import _lib
from . import plan as _plan

_lib.load_torch_ops()


def create_sequential_nn(layer_dims, activation=torch.nn.Tanh()):
    r"""Construct a feedforward PyTorch neural network (reference molann/ann.py:37-67).

    Children are named ``'%dth_layer'`` (Linear) and ``'activation of %dth_layer'`` (the SAME activation
    instance after every hidden layer); the last layer is linear.

    :raises AssertionError: if length of **layer_dims** is not larger than 1.
    """
    assert len(layer_dims) >= 2, \
        'Error: at least 2 layers are needed to define a neural network (length={})!'.format(len(layer_dims))
    layers = torch.nn.Sequential()
    n_linear = len(layer_dims) - 1
    for i in range(n_linear):
        layers.add_module('%dth_layer' % (i + 1), torch.nn.Linear(layer_dims[i], layer_dims[i + 1]))
        if i + 1 < n_linear:
            layers.add_module('activation of %dth_layer' % (i + 1), activation)
    return layers


def _int_buffer(values, shape=None):
    t = torch.tensor(list(values), dtype=torch.int32)
    if shape is not None:
        t = t.reshape(shape)
    return t


class AlignmentLayer(torch.nn.Module):
    r"""Optimal rigid alignment onto a reference (Kabsch), reference molann/ann.py:69-199.

    ``x -> (x - c(x)) R(x)`` with ``c`` the centroid of the alignment atoms and ``R`` the proper rotation
    that best superposes them onto the centred reference ``ref_x``.  Computed by the fused CUDA kernel
    ``molann_b200::align`` (quaternion eigenproblem instead of an SVD; same ``R``).

    Attributes: ``align_atom_indices``, ``input_atom_indices`` (0-based global), ``input_atom_num``,
    buffer ``ref_x`` (centred reference, fp32 ``[n_a, 3]``).

    Raises:
        ValueError: if some reference atom is not in ``input_atom_group``.
    """

    def __init__(self, align_atom_group, input_atom_group):
        super(AlignmentLayer, self).__init__()
        self.align_atom_indices = align_atom_group.ix.tolist()
        self.input_atom_indices = input_atom_group.ix.tolist()
        self.input_atom_num = len(input_atom_group)

        ref_x = torch.from_numpy(align_atom_group.positions)
        self.register_buffer('ref_x', ref_x)
        self.ref_x = self.ref_x - torch.mean(self.ref_x, 0)        # centred, still the registered buffer

        self._local_align_atom_indices = _plan.local_indices(
            self.align_atom_indices, self.input_atom_indices, "Atoms used for alignment must be among the input")
        self.register_buffer('_align_idx', _int_buffer(self._local_align_atom_indices), persistent=False)

    def show_info(self):
        """display indices of input atoms, indices and positions of the reference atoms used for alignment"""
        print(f'\n{self.input_atom_num} atoms used for input, (0-based) global indices: \n', self.input_atom_indices)
        print(f'\n{len(self._local_align_atom_indices)} atoms used for alignment, with (0-based) global indices: \n',
              self.align_atom_indices)
        print('local indices\n', self._local_align_atom_indices)
        print('\ncoordinates of reference state used in aligment:\n', self.ref_x.cpu().numpy())

    def forward(self, x):
        """align states ``x`` ([l, n_inp, 3]) by translation and rotation; same shape is returned.

        Raises:
            AssertionError: if `x` is not a Torch tensor with sizes :math:`[*, n_{inp},3]`.
        """
        assert isinstance(x, torch.Tensor), 'Input x is not a torch tensor'
        assert x.size(1) == self.input_atom_num and x.size(2) == 3, \
            f'Input should be a 3d torch tensor, with sizes [*, {self.input_atom_num}, 3]. Actual sizes: {x.shape}'
        return torch.ops.molann_b200.align(x, self._align_idx, self.ref_x)


class FeatureMap(torch.nn.Module):
    r"""Map coordinates to one feature (reference molann/ann.py:201-356).

    Output ``[l, d]`` with ``d = dim()``: angle -> cos (or radians), bond -> length, dihedral ->
    [cos, sin] (or radians in [-pi, pi]), position -> xyz of the feature's atoms.

    Raises:
        ValueError: if some atom used to define feature is not in the atom group for input.
    """

    def __init__(self, feature, input_atom_group, use_angle_value=False):
        super(FeatureMap, self).__init__()
        self.feature = feature
        self.type_id = feature.get_type_id()
        self.use_angle_value = use_angle_value
        self.input_atom_indices = input_atom_group.ix.tolist()
        self.input_atom_num = len(input_atom_group)

        atom_indices = feature.get_atom_indices() - 1               # 1-based -> 0-based (reference :258)
        self._local_atom_indices = _plan.local_indices(
            atom_indices, self.input_atom_indices, "Atoms used in feature must be among the input")
        entries, dim = _plan.compile_entries(self.type_id, self._local_atom_indices, bool(use_angle_value))
        self._dim = int(dim)
        self.register_buffer('_entries', torch.from_numpy(entries), persistent=False)
        self.register_buffer('_no_idx', torch.zeros(0, dtype=torch.int32), persistent=False)
        self.register_buffer('_no_ref', torch.zeros(0, 3, dtype=torch.float32), persistent=False)

    def dim(self):
        r"""total dimension of the feature (1, 1 or 2, 3n)."""
        output_dim = 0
        if self.type_id == 0 or self.type_id == 1:
            output_dim = 1
        if self.type_id == 2:
            if self.use_angle_value == True:
                output_dim = 1
            else:
                output_dim = 2
        if self.type_id == 3:
            output_dim = 3 * len(self.feature.get_atom_indices())
        return output_dim

    def forward(self, x):
        r"""map positions ``[l, n_inp, 3]`` to the feature ``[l, d]``.

        Raises:
            AssertionError: if `x` is not a Torch tensor with sizes :math:`[*, n_{inp},3]`.
        """
        assert isinstance(x, torch.Tensor), 'Input x is not a torch tensor'
        assert x.size(1) == self.input_atom_num and x.size(2) == 3, \
            f'Input should be a 3d torch tensor, with sizes [*, {self.input_atom_num}, 3]. Actual sizes: {x.shape}'
        return torch.ops.molann_b200.preprocess(x, self._no_idx, self._no_ref, self._entries, self._dim,
                                                self.use_angle_value)


class FeatureLayer(torch.nn.Module):
    r"""Map coordinates to all features of a list, concatenated in list order (reference :358-474).

    Raises:
        AssertionError: if feature_list is empty.
    """

    def __init__(self, feature_list, input_atom_group, use_angle_value=False):
        super(FeatureLayer, self).__init__()
        assert len(feature_list) > 0, 'Error: feature list is empty!'
        self.feature_list = feature_list
        self.feature_map_list = torch.nn.ModuleList(
            [FeatureMap(f, input_atom_group, use_angle_value) for f in feature_list])
        self.input_atom_num = len(input_atom_group)
        self.use_angle_value = bool(use_angle_value)

        entries, dim = _plan.compile_feature_program(
            [(fm.type_id, fm._local_atom_indices) for fm in self.feature_map_list], bool(use_angle_value))
        self._dim = int(dim)
        self.register_buffer('_entries', torch.from_numpy(entries), persistent=False)
        self.register_buffer('_no_idx', torch.zeros(0, dtype=torch.int32), persistent=False)
        self.register_buffer('_no_ref', torch.zeros(0, 3, dtype=torch.float32), persistent=False)

    def get_feature_info(self):
        r""":class:`pandas.DataFrame` with one row per feature"""
        return pd.concat([f.get_feature_info() for f in self.feature_list], ignore_index=True)

    def get_feature(self, idx):
        r"""the ``idx``-th :class:`molann_b200.feature.Feature` of the list"""
        return self.feature_list[idx]

    def output_dimension(self):
        r"""total dimension of the features in the list"""
        return sum([f_map.dim() for f_map in self.feature_map_list])

    def forward(self, x):
        """all features of states ``x`` ([l, n_inp, 3]) as columns ``[l, d]`` -- one fused kernel."""
        assert isinstance(x, torch.Tensor), 'Input x is not a torch tensor'
        assert x.size(1) == self.input_atom_num and x.size(2) == 3, \
            f'Input should be a 3d torch tensor, with sizes [*, {self.input_atom_num}, 3]. Actual sizes: {x.shape}'
        return torch.ops.molann_b200.preprocess(x, self._no_idx, self._no_ref, self._entries, self._dim,
                                                self.use_angle_value)


class PreprocessingANN(torch.nn.Module):
    """Alignment (optional) followed by the feature layer (reference molann/ann.py:476-565).

    With an :class:`AlignmentLayer` and a :class:`FeatureLayer` the two run as ONE kernel: bond / angle /
    dihedral features are rigid-motion invariant and are evaluated on the raw coordinates, only position
    features go through the optimal rotation; the aligned copy of ``x`` is never materialised.
    """
    __constants__ = ['_mode', '_fused_align']

    def __init__(self, align_layer, feature_layer):
        super(PreprocessingANN, self).__init__()
        if align_layer is not None:
            self.align_layer = align_layer
        else:
            self.align_layer = torch.nn.Identity()
        self.feature_layer = feature_layer
        # 2: fused align+features, 1: features only (no alignment), 0: generic composition
        if isinstance(feature_layer, FeatureLayer) and isinstance(self.align_layer, AlignmentLayer):
            self._mode = 2
        elif isinstance(feature_layer, FeatureLayer) and isinstance(self.align_layer, torch.nn.Identity):
            self._mode = 1
        else:
            self._mode = 0
        self._fused_align = (self._mode == 2)      # bool constants give TorchScript a static `if`

    def output_dimension(self):
        """the dimension of the output layer"""
        return self.feature_layer.output_dimension()

    def forward(self, x):
        """align states and map them to features: ``[l, n_inp, 3] -> [l, d]``"""
        if self._fused_align:
            assert isinstance(x, torch.Tensor), 'Input x is not a torch tensor'
            assert x.size(1) == self.feature_layer.input_atom_num and x.size(2) == 3, \
                f'Input should be a 3d torch tensor, with sizes [*, {self.feature_layer.input_atom_num}, 3]. Actual sizes: {x.shape}'
            return torch.ops.molann_b200.preprocess(x, self.align_layer._align_idx, self.align_layer.ref_x,
                                                    self.feature_layer._entries, self.feature_layer._dim,
                                                    self.feature_layer.use_angle_value)
        else:
            return self.feature_layer(self.align_layer(x))


def _fusable_activation(ann_layers):
    """-> activation id if ``ann_layers`` is Sequential(Linear, act, Linear, ..., Linear), else -1."""
    if not isinstance(ann_layers, torch.nn.Sequential):
        return -1
    mods = list(ann_layers._modules.values())           # not children(): the shared activation repeats
    if len(mods) == 0 or len(mods) % 2 == 0:
        return -1
    act_name = None
    for i, m in enumerate(mods):
        if i % 2 == 0:
            if type(m) is not torch.nn.Linear or m.bias is None:
                return -1
        else:
            name = type(m).__name__
            if type(m) not in (torch.nn.Tanh, torch.nn.ReLU, torch.nn.Sigmoid, torch.nn.Identity):
                return -1
            if act_name is not None and name != act_name:
                return -1
            act_name = name
    if (len(mods) + 1) // 2 > _lib.MAX_LAYERS:
        return -1
    return _plan.ACT_IDS[act_name] if act_name is not None else _plan.ACT_IDS["Identity"]


class MolANN(torch.nn.Module):
    """Preprocessing layer followed by trainable layers (reference molann/ann.py:567-624).

    When ``ann_layers`` is a ``create_sequential_nn``-style ``Sequential`` (Linear / shared activation
    in {Tanh, ReLU, Sigmoid, Identity}) on top of a :class:`PreprocessingANN`, the whole model -- align,
    features, MLP -- and its backward to coordinates run as ONE fused kernel each
    (``molann_b200::molann``).  Any other ``ann_layers`` module is composed with the fused preprocessing
    op on the same device.
    """
    __constants__ = ['_mode', '_act_id', '_fused', '_fused_align']

    def __init__(self, preprocessing_layer, ann_layers):
        super(MolANN, self).__init__()
        self.preprocessing_layer = preprocessing_layer
        self.ann_layers = ann_layers
        act = _fusable_activation(ann_layers)
        pp_mode = preprocessing_layer._mode if isinstance(preprocessing_layer, PreprocessingANN) else 0
        if act >= 0 and pp_mode == 2:
            self._mode = 2
        elif act >= 0 and pp_mode == 1:
            self._mode = 1
        else:
            self._mode = 0
        self._act_id = act if act >= 0 else 0
        self._fused = (self._mode != 0)            # bool constants give TorchScript a static `if`
        self._fused_align = (self._mode == 2)
        self._n_mlp_modules = len(ann_layers._modules) if isinstance(ann_layers, torch.nn.Sequential) else -1

    @torch.jit.unused
    def _mlp_unchanged(self) -> bool:
        """Eager-mode guard: the fused kernel was chosen for the ``ann_layers`` seen at construction.  If the user
        has since added / swapped a child (another activation, an extra layer), fall back to the composition the
        reference always evaluates instead of silently running the stale structure.  (A scripted module is frozen,
        so the check is not needed -- nor possible -- there.)"""
        if len(self.ann_layers._modules) != self._n_mlp_modules:
            return False
        return _fusable_activation(self.ann_layers) == self._act_id

    def get_preprocessing_layer(self):
        """the :class:`PreprocessingANN` of this model"""
        return self.preprocessing_layer

    @torch.jit.unused
    def _value_and_jacobian_composed(self, x):
        xg = x.detach().requires_grad_(True)
        y = self.ann_layers(self.preprocessing_layer(xg))
        planes = [torch.autograd.grad(y[:, o].sum(), xg, retain_graph=True)[0] for o in range(y.shape[1])]
        return y.detach(), torch.stack(planes).detach()

    @torch.jit.unused
    def _value_and_grad_composed(self, x, cotangent):
        xg = x.detach().requires_grad_(True)
        y = self.ann_layers(self.preprocessing_layer(xg))
        (g,) = torch.autograd.grad(y, xg, cotangent)
        return y.detach(), g.detach()

    @torch.jit.export
    def value_and_grad(self, x, cotangent):
        """``(y, d<cotangent, y>/dx)`` in ONE fused pass over ``x`` (the biasing-force path of MD plugins).

        Not part of the reference API: it replaces ``y = model(x); torch.autograd.grad(y, x, cotangent)``
        (two passes over the coordinates) when no autograd graph is needed.  ``cotangent``: ``[l, k]``.
        """
        assert x.size(1) == self.preprocessing_layer.feature_layer.input_atom_num and x.size(2) == 3, \
            'Input should be a 3d torch tensor with sizes [*, n_inp, 3]'
        if not torch.jit.is_scripting():
            if self._fused and not self._mlp_unchanged():
                return self._value_and_grad_composed(x, cotangent)
        if not self._fused:
            xg = x.detach().requires_grad_(True)
            y = self.forward(xg)
            cots: List[Optional[torch.Tensor]] = [cotangent]
            g = torch.autograd.grad([y], [xg], cots)[0]
            assert g is not None
            return y.detach(), g
        else:
            flayer = self.preprocessing_layer.feature_layer
            params: List[torch.Tensor] = []
            for layer in self.ann_layers:
                if hasattr(layer, 'weight'):
                    params.append(layer.weight)
                    params.append(layer.bias)
            if self._fused_align:
                return torch.ops.molann_b200.value_and_grad(
                    x, cotangent, self.preprocessing_layer.align_layer._align_idx,
                    self.preprocessing_layer.align_layer.ref_x, flayer._entries, flayer._dim,
                    flayer.use_angle_value, params, self._act_id)
            else:
                return torch.ops.molann_b200.value_and_grad(
                    x, cotangent, flayer._no_idx, flayer._no_ref, flayer._entries, flayer._dim,
                    flayer.use_angle_value, params, self._act_id)

    @torch.jit.export
    def value_and_jacobian(self, x):
        """``(y, J)`` with ``J[o] = d y[:, o] / dx``, shape ``[k, l, n_inp, 3]``: the force on every atom from every
        output (collective variable) in ONE pass over ``x`` -- what an MD plugin that biases several CVs needs per
        step.  Not part of the reference API: it replaces ``k`` calls of ``torch.autograd.grad(y[:, o].sum(), x)``."""
        assert x.size(1) == self.preprocessing_layer.feature_layer.input_atom_num and x.size(2) == 3, \
            'Input should be a 3d torch tensor with sizes [*, n_inp, 3]'
        if not torch.jit.is_scripting():
            if self._fused and not self._mlp_unchanged():
                return self._value_and_jacobian_composed(x)
        if not self._fused:
            y0 = self.forward(x.detach())
            planes: List[torch.Tensor] = []
            for o in range(y0.size(1)):
                cot = torch.zeros_like(y0)
                cot[:, o] = 1.0
                planes.append(self.value_and_grad(x, cot)[1])
            return y0.detach(), torch.stack(planes)
        else:
            flayer = self.preprocessing_layer.feature_layer
            params: List[torch.Tensor] = []
            for layer in self.ann_layers:
                if hasattr(layer, 'weight'):
                    params.append(layer.weight)
                    params.append(layer.bias)
            if self._fused_align:
                return torch.ops.molann_b200.value_and_jacobian(
                    x, self.preprocessing_layer.align_layer._align_idx, self.preprocessing_layer.align_layer.ref_x,
                    flayer._entries, flayer._dim, flayer.use_angle_value, params, self._act_id)
            else:
                return torch.ops.molann_b200.value_and_jacobian(
                    x, flayer._no_idx, flayer._no_ref, flayer._entries, flayer._dim, flayer.use_angle_value,
                    params, self._act_id)

    def forward(self, x):
        """the forward map ``[l, n_inp, 3] -> [l, k]``"""
        if not torch.jit.is_scripting():
            if self._fused and not self._mlp_unchanged():
                return self.ann_layers(self.preprocessing_layer(x))
        if not self._fused:
            return self.ann_layers(self.preprocessing_layer(x))
        else:
            flayer = self.preprocessing_layer.feature_layer
            assert isinstance(x, torch.Tensor), 'Input x is not a torch tensor'
            assert x.size(1) == flayer.input_atom_num and x.size(2) == 3, \
                f'Input should be a 3d torch tensor, with sizes [*, {flayer.input_atom_num}, 3]. Actual sizes: {x.shape}'
            params: List[torch.Tensor] = []
            for layer in self.ann_layers:
                if hasattr(layer, 'weight'):
                    params.append(layer.weight)
                    params.append(layer.bias)
            if self._fused_align:
                return torch.ops.molann_b200.molann(
                    x, self.preprocessing_layer.align_layer._align_idx, self.preprocessing_layer.align_layer.ref_x,
                    flayer._entries, flayer._dim, flayer.use_angle_value, params, self._act_id)
            else:
                return torch.ops.molann_b200.molann(
                    x, flayer._no_idx, flayer._no_ref, flayer._entries, flayer._dim, flayer.use_angle_value,
                    params, self._act_id)
