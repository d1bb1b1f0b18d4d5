"""Generate tests/golden/*.npz by running the UNMODIFIED reference (zwpku/molann) on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (``/root/reference`` must exist):

    python -m oracle.make_golden

The reference's own tests store no expected values (test/test_molann.py has no assert), so these files
ARE the pinned known answers: outputs of the reference's fp32 and fp64 (``module.double()``) forward and
autograd on seeded inputs, plus its integer index maps.  MDAnalysis is replaced by the duck-typed
AtomGroup of molann_b200/atomgroup.py.
"""
import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from molann_b200 import synthetic as S                    # noqa: E402
from molann_b200.atomgroup import Universe                # noqa: E402
from oracle.ref_loader import load_reference              # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def ref_api():
    ref = load_reference()
    if ref is None:
        raise SystemExit("reference not found (/root/reference or baseline/_ref)")
    ann, feature, root = ref
    return SimpleNamespace(Feature=feature.Feature, FeatureLayer=ann.FeatureLayer, FeatureMap=ann.FeatureMap,
                           AlignmentLayer=ann.AlignmentLayer, PreprocessingANN=ann.PreprocessingANN,
                           MolANN=ann.MolANN, create_sequential_nn=ann.create_sequential_nn), root


def per_frame(model, x):
    """Evaluate frame by frame (avoids the reference's L == 3 torch.cross quirk, App. B #1)."""
    return torch.cat([model(x[i:i + 1]) for i in range(x.shape[0])], dim=0)


def run_model(model, x, cot):
    out = {}
    for tag, dt in (("32", torch.float32), ("64", torch.float64)):
        m = model.double() if dt == torch.float64 else model.float()
        xx = x.to(dt).clone().requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            y = m(xx)
            (gx,) = torch.autograd.grad((y * cot.to(dt)).sum(), xx)
        out["y" + tag] = y.detach().numpy()
        out["gx" + tag] = gx.numpy()
    model.float()
    return out


def config_case(api, name, L, seed):
    spec = S.get_spec(name)
    model, _ = S.build_model(spec, api)
    x = S.make_frames(spec, L, seed=seed)
    g = torch.Generator().manual_seed(seed + 7)
    cot = torch.randn(L, spec.out_dim(), generator=g)
    out = {"x": x.numpy(), "cot": cot.numpy()}
    out.update(run_model(model, x, cot))
    pp = model.get_preprocessing_layer()
    gf = torch.Generator().manual_seed(seed + 11)
    cotf = torch.randn(L, spec.feature_dim(), generator=gf)
    res = run_model(pp, x, cotf)
    out.update({"cotf": cotf.numpy(), "feat32": res["y32"], "feat64": res["y64"],
                "gxf32": res["gx32"], "gxf64": res["gx64"]})
    for k, v in model.float().state_dict().items():
        out["sd::" + k] = v.numpy()
    # parameter gradients (fp64) for the training path
    m64 = model.double()
    xx = x.double()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        loss = (m64(xx) * cot.double()).sum()
    grads = torch.autograd.grad(loss, list(m64.parameters()))
    for (k, _), gp in zip(m64.named_parameters(), grads):
        out["gp64::" + k] = gp.numpy()
    model.float()
    return out


def fixture_cases(api):
    """SURVEY App. C known answers on the 22-atom fixture (single frame)."""
    pos = S.ala2_positions()
    u = Universe(pos)
    x = torch.from_numpy(pos).unsqueeze(0)
    out = {"x": pos}
    one = lambda *ids: u.select_ix([i - 1 for i in ids])       # 1-based, order preserving
    hist = [("d1", "dihedral", (5, 7, 9, 15)), ("d2", "dihedral", (7, 9, 15, 17)), ("b1", "bond", (2, 5)),
            ("b2", "bond", (5, 6)), ("a1", "angle", (20, 19, 21)), ("a2", "angle", (16, 15, 17))]
    feats = [api.Feature(n, t, one(*ids)) for n, t, ids in hist]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["hist_cs"] = api.FeatureLayer(feats, u.atoms, use_angle_value=False)(x).numpy()
        out["hist_angle"] = api.FeatureLayer(feats, u.atoms, use_angle_value=True)(x).numpy()
        extra = [(1, 2, 5, 6), (3, 2, 5, 7), (10, 9, 11, 12), (9, 15, 17, 19)]
        fe = [api.Feature("e%d" % i, "dihedral", one(*ids)) for i, ids in enumerate(extra)]
        out["extra_dihedrals"] = api.FeatureLayer(fe, u.atoms, use_angle_value=False)(x).numpy()
        # test_FeatureLayer (test/test_molann.py:50-57): input atoms 1-5; select_atoms sorts
        inp = one(1, 2, 3, 4, 5)
        fl = api.FeatureLayer([api.Feature("n", "dihedral", one(1, 2, 3, 4)), api.Feature("n", "bond", one(1, 3)),
                               api.Feature("n", "angle", one(1, 2, 3))], inp, use_angle_value=False)
        out["test_feature_layer"] = fl(torch.from_numpy(inp.positions).unsqueeze(0)).numpy()
        # AlignmentLayer(atoms 1,2,5; all 22 as input) (test/test_molann.py:39-44)
        al = api.AlignmentLayer(one(1, 2, 5), u.atoms)
        out["align_ref_x"] = al.ref_x.numpy()
        out["align_self"] = al(x).numpy()
        a, b = 0.7, 1.1
        Rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
        Rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
        moved = (pos.astype(np.float64) @ Rz @ Rx + np.array([1.5, -2.25, 3.0])).astype(np.float32)
        out["align_moved_x"] = moved
        out["align_moved"] = al(torch.from_numpy(moved).unsqueeze(0)).numpy()
    return out


def index_cases(api):
    """Integer index maps of the reference for permuted / subset input groups (bit-exact contract)."""
    rng = np.random.RandomState(1234)
    pos = S.ala2_positions()
    u = Universe(pos)
    out = {}
    for c in range(6):
        n_in = int(rng.randint(8, 23))
        inp_ix = rng.permutation(22)[:n_in]
        inp = u.select_ix(inp_ix)
        al_ix = rng.permutation(inp_ix)[:int(rng.randint(3, 7))]
        al = api.AlignmentLayer(u.select_ix(al_ix), inp)
        out["c%d_input" % c] = np.asarray(inp_ix, dtype=np.int64)
        out["c%d_align" % c] = np.asarray(al_ix, dtype=np.int64)
        out["c%d_align_local" % c] = np.asarray(al._local_align_atom_indices, dtype=np.int64)
        kinds = [("dihedral", 4), ("bond", 2), ("angle", 3), ("position", int(rng.randint(1, 5)))]
        feats, spec_rows = [], []
        for k, (tp, m) in enumerate(kinds):
            ix = rng.permutation(inp_ix)[:m]
            feats.append(api.Feature("f%d" % k, tp, u.select_ix(ix)))
            spec_rows.append(ix)
            out["c%d_f%d_atoms" % (c, k)] = np.asarray(ix, dtype=np.int64)
        for ua in (False, True):
            fl = api.FeatureLayer(feats, inp, use_angle_value=ua)
            out["c%d_dims_ua%d" % (c, int(ua))] = np.asarray([fm.dim() for fm in fl.feature_map_list], dtype=np.int64)
            out["c%d_outdim_ua%d" % (c, int(ua))] = np.asarray([fl.output_dimension()], dtype=np.int64)
        for k, fm in enumerate(fl.feature_map_list):
            out["c%d_f%d_local" % (c, k)] = np.asarray(fm._local_atom_indices, dtype=np.int64)
            out["c%d_f%d_type" % (c, k)] = np.asarray([fm.type_id], dtype=np.int64)
    return out


def main():
    api, root = ref_api()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(4)
    np.savez_compressed(os.path.join(GOLD, "fixture.npz"), **fixture_cases(api))
    np.savez_compressed(os.path.join(GOLD, "index_maps.npz"), **index_cases(api))
    for name, L, seed in (("C1", 96, 1101), ("C2", 192, 1202), ("C3s", 24, 1303)):
        np.savez_compressed(os.path.join(GOLD, "config_%s.npz" % name), **config_case(api, name, L, seed))
    with open(os.path.join(GOLD, "PROVENANCE.txt"), "w") as fh:
        fh.write("generated by oracle/make_golden.py from the unmodified reference at %s\n" % root)
        fh.write("torch %s, numpy %s\n" % (torch.__version__, np.__version__))
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
