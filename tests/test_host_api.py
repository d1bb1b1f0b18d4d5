"""Host-side logic of the drop-in API: construction, integer index programs (bit-exact vs the reference's
maps), error conventions, state_dict keys, TorchScript script/save/load, and the loud no-CPU behaviour."""
import io
import os

import numpy as np
import pandas as pd
import pytest
import torch

from helpers import S, golden, spec_program
from molann_b200 import plan as P
from molann_b200.ann import (AlignmentLayer, FeatureLayer, FeatureMap, MolANN, PreprocessingANN,
                             create_sequential_nn)
from molann_b200.atomgroup import AtomGroup, Universe
from molann_b200.feature import Feature, FeatureFileReader


@pytest.fixture(scope="module")
def u():
    return Universe(S.ala2_positions())


def test_import_paths_are_drop_in():
    import molann.ann as a
    import molann.feature as f
    assert a.MolANN is MolANN and f.Feature is Feature
    for name in ("Feature", "FeatureFileReader"):
        assert hasattr(f, name)
    for name in ("AlignmentLayer", "FeatureMap", "FeatureLayer", "PreprocessingANN", "MolANN", "create_sequential_nn"):
        assert hasattr(a, name)


def test_feature_validation(u):
    f = Feature("b", "bond", u.select_ix([0, 1]))
    assert (f.get_name(), f.get_type(), f.get_type_id()) == ("b", "bond", 1)
    assert f.get_atom_indices().tolist() == [1, 2]                 # 1-based (reference feature.py:123)
    info = f.get_feature_info()
    assert isinstance(info, pd.DataFrame)
    assert list(info.columns) == ["name", "type", "type_id", "atom indices (1-based)"]
    assert [Feature("x", t, u.select_ix(list(range(n)))).get_type_id()
            for t, n in (("angle", 3), ("bond", 2), ("dihedral", 4), ("position", 5))] == [0, 1, 2, 3]
    with pytest.raises(NotImplementedError):
        Feature("x", "torsion", u.select_ix([0, 1]))
    with pytest.raises(IndexError):
        Feature("x", "bond", u.select_ix([3, 3]))
    for t, n in (("angle", 2), ("bond", 3), ("dihedral", 3)):
        with pytest.raises(AssertionError):
            Feature("x", t, u.select_ix(list(range(n))))


def test_create_sequential_nn_layout():
    act = torch.nn.ReLU()
    net = create_sequential_nn([10, 5, 7, 1], act)
    assert list(net._modules.keys()) == ["1th_layer", "activation of 1th_layer", "2th_layer",
                                         "activation of 2th_layer", "3th_layer"]
    assert net._modules["activation of 1th_layer"] is act and net._modules["activation of 2th_layer"] is act
    assert [tuple(p.shape) for p in net.parameters()] == [(5, 10), (5,), (7, 5), (7,), (1, 7), (1,)]
    assert isinstance(create_sequential_nn([3, 2])._modules["1th_layer"], torch.nn.Linear)
    with pytest.raises(AssertionError):
        create_sequential_nn([10])


def test_index_programs_bit_exact_vs_reference_goldens(u):
    """Local index maps, dims and column offsets equal what the reference computes (ann.py:144,258-263,473)."""
    g = golden("index_maps")
    for c in range(6):
        inp = u.select_ix(g["c%d_input" % c])
        al = AlignmentLayer(u.select_ix(g["c%d_align" % c]), inp)
        assert al._local_align_atom_indices == g["c%d_align_local" % c].tolist()
        assert al._align_idx.dtype == torch.int32 and al._align_idx.tolist() == g["c%d_align_local" % c].tolist()
        assert al.align_atom_indices == g["c%d_align" % c].tolist()
        assert al.input_atom_indices == g["c%d_input" % c].tolist()
        kinds = ["dihedral", "bond", "angle", "position"]
        feats = [Feature("f%d" % k, kinds[k], u.select_ix(g["c%d_f%d_atoms" % (c, k)])) for k in range(4)]
        for ua in (False, True):
            fl = FeatureLayer(feats, inp, use_angle_value=ua)
            dims = [fm.dim() for fm in fl.feature_map_list]
            assert dims == g["c%d_dims_ua%d" % (c, int(ua))].tolist()
            assert fl.output_dimension() == int(g["c%d_outdim_ua%d" % (c, int(ua))][0]) == fl._dim
            ent = fl._entries.numpy()
            col, row = 0, 0
            for k, fm in enumerate(fl.feature_map_list):
                loc = g["c%d_f%d_local" % (c, k)].tolist()
                assert fm._local_atom_indices == loc and fm.type_id == int(g["c%d_f%d_type" % (c, k)][0])
                if fm.type_id == 3:
                    for j, a in enumerate(loc):
                        assert ent[row].tolist() == [3, a, 0, 0, 0, col + 3 * j]
                        row += 1
                else:
                    assert ent[row].tolist() == [fm.type_id] + loc + [0] * (4 - len(loc)) + [col]
                    row += 1
                col += dims[k]
            assert row == ent.shape[0] and ent.dtype == np.int32


def test_atoms_must_be_among_input(u):
    inp = u.select_ix([0, 1, 2, 3])
    with pytest.raises(ValueError, match="Atoms used for alignment must be among the input"):
        AlignmentLayer(u.select_ix([0, 9]), inp)
    with pytest.raises(ValueError, match="Atoms used in feature must be among the input"):
        FeatureMap(Feature("b", "bond", u.select_ix([0, 9])), inp)
    with pytest.raises(AssertionError):
        FeatureLayer([], inp)


def test_alignment_layer_buffers(u):
    g = golden("fixture")
    al = AlignmentLayer(u.select_ix([0, 1, 4]), u.atoms)
    np.testing.assert_allclose(al.ref_x.numpy(), g["align_ref_x"], atol=1e-7)
    assert al.ref_x.dtype == torch.float32 and al.input_atom_num == 22
    assert list(al.state_dict().keys()) == ["ref_x"]


@pytest.mark.parametrize("name", ["C1", "C2", "C3s"])
def test_state_dict_keys_match_reference(name):
    spec = S.get_spec(name)
    model, _ = S.build_model(spec)
    g = golden("config_" + name)
    ref_keys = [k[4:] for k in g.files if k.startswith("sd::")]
    assert list(model.state_dict().keys()) == ref_keys
    # a reference checkpoint loads with strict=True
    model.load_state_dict({k: torch.from_numpy(g["sd::" + k]) for k in ref_keys}, strict=True)
    assert model.get_preprocessing_layer() is model.preprocessing_layer
    assert model.preprocessing_layer.output_dimension() == spec.feature_dim()


def test_preprocessing_none_is_identity(u):
    fl = FeatureLayer([Feature("d", "dihedral", u.select_ix([0, 1, 2, 3]))], u.atoms)
    pp = PreprocessingANN(None, fl)
    assert isinstance(pp.align_layer, torch.nn.Identity) and pp.output_dimension() == 2


@pytest.mark.parametrize("name", ["C1", "C2"])
def test_torchscript_script_save_load_cpu(name, tmp_path):
    """Scripting needs no GPU: the archive must carry the op call, the program buffers and the weights."""
    spec = S.get_spec(name)
    model, _ = S.build_model(spec)
    scripted = torch.jit.script(model)
    kinds = [n.kind() for n in scripted.graph.nodes()]
    assert "molann_b200::molann" in kinds
    path = os.path.join(tmp_path, "model.pt")
    scripted.save(path)
    loaded = torch.jit.load(path)
    n_lin = len(spec.layer_dims) - 1
    code = loaded.code
    assert code.count("torch.append(params") == 2 * n_lin           # every Linear collected (App. B #11)
    sd = loaded.state_dict()
    for k, v in model.state_dict().items():
        assert torch.equal(sd[k], v)
    ent = sd["preprocessing_layer.feature_layer._entries"]
    assert ent.dtype == torch.int32 and torch.equal(ent, model.preprocessing_layer.feature_layer._entries)
    for part in (model.preprocessing_layer, model.preprocessing_layer.feature_layer,
                 model.preprocessing_layer.feature_layer.feature_map_list[0]):
        torch.jit.script(part)


def test_generic_ann_layers_compose():
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)

    class Squared(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(30, 2)

        def forward(self, f):
            return self.lin(f) ** 2
    m = MolANN(model.preprocessing_layer, Squared())
    assert m._fused is False
    torch.jit.script(m)
    assert model._fused is True and model._fused_align is True


def test_cpu_input_raises_loudly():
    """No CPU fallback: a CPU tensor must fail with a clear message, never compute."""
    spec = S.get_spec("C2")
    model, _ = S.build_model(spec)
    x = S.make_frames(spec, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.preprocessing_layer(x)
    with pytest.raises(RuntimeError, match="CUDA"):
        model.preprocessing_layer.align_layer(x)
    with pytest.raises(AssertionError):
        model(torch.zeros(4, 21, 3))
    with pytest.raises(AssertionError):
        model.preprocessing_layer.feature_layer(np.zeros((4, 22, 3)))


def test_pdb_parser_and_feature_file_reader(tmp_path):
    pdb = S.write_ala2_pdb(os.path.join(tmp_path, "ala2.pdb"))
    uni = Universe(pdb)
    assert len(uni.atoms) == 22
    np.testing.assert_allclose(uni.atoms.positions, S.ala2_positions(), atol=1e-6)
    assert uni.select_atoms("bynum 1 3 2 4").ix.tolist() == [0, 1, 2, 3]        # sorted, like MDAnalysis
    assert (uni.select_atoms("bynum 5") + uni.select_atoms("bynum 2")).ix.tolist() == [4, 1]
    assert uni.select_atoms("resid 2").ix.tolist() == list(range(6, 16))
    assert uni.select_atoms("heavy").ix.tolist() == S.ALA2_HEAVY
    ffile = os.path.join(tmp_path, "feature.txt")
    with open(ffile, "w") as fh:
        fh.write("# comment\n\n[Preprocessing]\np1, position, resid 2\n[End]\n[Histogram]\n"
                 "d1, dihedral, bynum 5, bynum 7, bynum 9, bynum 15\nb1, bond, bynum 2 5\n"
                 "a1, angle, bynum 20, bynum 19, bynum 21\n[End]\n[Output]\nd1, dihedral, bynum 5 7 9 15\n[End]\n")
    rd = FeatureFileReader(ffile, "Histogram", uni)
    feats = rd.read()
    assert [f.get_name() for f in feats] == ["d1", "b1", "a1"] and rd.get_num_of_features() == 3
    assert feats[2].get_atom_indices().tolist() == [20, 19, 21]
    assert len(rd.get_feature_info()) == 3 and rd.get_feature_list() is feats
    pre = FeatureFileReader(ffile, "Preprocessing", uni).read()
    assert pre[0].get_type() == "position" and len(pre[0].atom_group) == 10


def test_atomgroup_protocol():
    ag = AtomGroup([3, 1], np.arange(15, dtype=np.float32).reshape(5, 3))
    assert ag.ix.tolist() == [3, 1] and len(ag) == 2 and len(set(ag)) == 2
    assert ag.positions.dtype == np.float32 and ag.positions[0].tolist() == [9, 10, 11]
    with pytest.raises(IndexError):
        AtomGroup([7], np.zeros((5, 3)))


def test_numa_binding_helper_is_harmless_without_a_gpu():
    """bench.py binds each rank to the CPUs next to its GPU before it pins host buffers; on a box without NVML /
    GPUs the helper must change nothing and report 0."""
    import os
    from molann_b200.stream import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    kept = bind_to_gpu_numa_node(0)
    assert kept == 0 or kept == len(os.sched_getaffinity(0))
    if kept == 0:
        assert os.sched_getaffinity(0) == before


def test_autoencoder_step_surface():
    """Host logic of the C4 trainer on plain CPU modules: loss normalised by the GLOBAL batch, SGD update applied."""
    import torch
    from molann_b200.train import AutoencoderStep

    class Enc(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lin = torch.nn.Linear(6, 2)

        def get_preprocessing_layer(self):
            return lambda x: x.reshape(x.shape[0], -1)

        def forward(self, x):
            return self.lin(x.reshape(x.shape[0], -1))

    torch.manual_seed(0)
    enc, dec = Enc().double(), torch.nn.Linear(2, 6).double()
    x = torch.randn(32, 2, 3, dtype=torch.float64)
    tr = AutoencoderStep(enc, dec, lr=0.1, global_frames=64)
    w0 = dec.weight.detach().clone()
    loss = tr.loss_and_grads(x)
    want = ((dec(enc(x)) - x.reshape(32, -1)) ** 2).sum() / (64 * 6)
    assert abs(float(loss) - float(want)) < 1e-12
    g = dec.weight.grad.clone()
    tr.step(x)
    assert torch.allclose(dec.weight.detach(), w0 - 0.1 * g, atol=1e-12)


def test_wire_format_host_codec():
    """int16 wire format (molann_b200.stream): coding error <= resolution / 2, exact round trip of decoded frames."""
    import torch
    from molann_b200 import synthetic as S
    from molann_b200.stream import dequantize_frames, quantize_frames
    x = S.make_frames(S.get_spec("C2"), 4000)
    q, origin, res = quantize_frames(x)
    assert q.dtype == torch.int16 and tuple(q.shape) == tuple(x.shape) and int(q.abs().max()) <= 32767
    xd = dequantize_frames(q, origin, res)
    assert float((xd - x).abs().max()) <= 0.5 * res + 4e-7 * float(x.abs().max())   # + fp32 rounding of the sum
    q2, o2, r2 = quantize_frames(x, resolution=0.01)
    assert abs(r2 - 0.01) < 1e-9 and float((dequantize_frames(q2, o2, r2) - x).abs().max()) <= 0.5 * 0.01 + 4e-7 * float(x.abs().max())
    # the decoded frames are fixed points of the codec at the same origin / resolution
    q3 = torch.round((xd.double() - torch.tensor(origin).double()) / float(torch.tensor(res, dtype=torch.float32)))
    assert torch.equal(q3.to(torch.int16), q)
