"""Deterministic synthetic systems and trajectories for the BASELINE.json configurations.

C1..C5 follow SURVEY.md section 8(d).  A ``SystemSpec`` is API-agnostic: ``build_model`` takes a
namespace that provides ``Feature, FeatureLayer, AlignmentLayer, PreprocessingANN, MolANN,
create_sequential_nn`` -- this repo's ``molann_b200.ann``/``feature`` or (in tests / the CPU
baseline) the reference's modules -- so both sides are constructed from identical inputs.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import List, Optional, Tuple

import numpy as np
import torch

from .atomgroup import Universe

# Alanine dipeptide (ACE-ALA-NME) in vacuum, 22 atoms: (atom name, residue id, x, y, z) in Angstrom.
# Same system as the reference's test fixture test/alanine-dipeptide-vacuum.pdb:2-23.
ALA2_ATOMS = [
    ("1HH3", 1, 2.000, 1.000, -0.000), ("CH3", 1, 2.000, 2.090, 0.000),
    ("2HH3", 1, 1.486, 2.454, 0.890), ("3HH3", 1, 1.486, 2.454, -0.890),
    ("C", 1, 3.427, 2.641, -0.000), ("O", 1, 4.391, 1.877, -0.000),
    ("N", 2, 3.555, 3.970, -0.000), ("H", 2, 2.733, 4.556, -0.000),
    ("CA", 2, 4.853, 4.614, -0.000), ("HA", 2, 5.408, 4.316, 0.890),
    ("CB", 2, 5.661, 4.221, -1.232), ("1HB", 2, 5.123, 4.521, -2.131),
    ("2HB", 2, 6.630, 4.719, -1.206), ("3HB", 2, 5.809, 3.141, -1.241),
    ("C", 2, 4.713, 6.129, 0.000), ("O", 2, 3.601, 6.653, 0.000),
    ("N", 3, 5.846, 6.835, 0.000), ("H", 3, 6.737, 6.359, -0.000),
    ("CH3", 3, 5.846, 8.284, 0.000), ("1HH3", 3, 4.819, 8.648, 0.000),
    ("2HH3", 3, 6.360, 8.648, 0.890), ("3HH3", 3, 6.360, 8.648, -0.890),
]
ALA2_RESNAMES = {1: "ACE", 2: "ALA", 3: "NME"}
ALA2_HEAVY = [1, 4, 5, 6, 8, 10, 14, 15, 16, 18]     # 0-based heavy atoms


def ala2_positions() -> np.ndarray:
    return np.asarray([[a[2], a[3], a[4]] for a in ALA2_ATOMS], dtype=np.float32)


def write_ala2_pdb(path: str) -> str:
    """Write the 22-atom system as a fixed-column PDB file (for the PDB parser / Universe)."""
    with open(path, "w") as fh:
        fh.write("REMARK  ACE\n")
        for i, (name, resid, x, y, z) in enumerate(ALA2_ATOMS):
            nm = name if len(name) == 4 else " " + name.ljust(3)
            fh.write("ATOM  %5d %4s %3s  %4d    %8.3f%8.3f%8.3f\n"
                     % (i + 1, nm, ALA2_RESNAMES[resid], resid, x, y, z))
        fh.write("TER\nEND\n")
    return path


def chain_positions(n: int, seed: int, bond: float = 1.5, angle_deg: float = 111.0) -> np.ndarray:
    """Freely-rotating chain: fixed bond length / bond angle, torsions U(-pi, pi)."""
    rng = np.random.RandomState(seed)
    theta = math.radians(angle_deg)
    pos = np.zeros((n, 3), dtype=np.float64)
    pos[1] = [bond, 0.0, 0.0]
    pos[2] = pos[1] + bond * np.array([-math.cos(theta), math.sin(theta), 0.0])
    tors = rng.uniform(-math.pi, math.pi, size=n)
    for i in range(3, n):
        a, b, c = pos[i - 3], pos[i - 2], pos[i - 1]
        bc = c - b
        bc /= np.linalg.norm(bc)
        nrm = np.cross(b - a, bc)
        nrm /= np.linalg.norm(nrm)
        m = np.cross(nrm, bc)
        d2 = np.array([-bond * math.cos(theta),
                       bond * math.sin(theta) * math.cos(tors[i]),
                       bond * math.sin(theta) * math.sin(tors[i])])
        pos[i] = c + d2[0] * bc + d2[1] * m + d2[2] * nrm
    return pos.astype(np.float32)


@dataclass
class SystemSpec:
    name: str
    positions: np.ndarray                         # [n_total, 3] reference structure
    input_ix: List[int]                           # 0-based global indices of the input atom group
    align_ix: Optional[List[int]]                 # 0-based global indices (None: no alignment)
    features: List[Tuple[str, str, List[int]]]    # (name, type, 0-based global indices in order)
    use_angle_value: bool
    layer_dims: Optional[List[int]]               # None: preprocessing only
    noise: float
    trans_sigma: float
    rotate: bool
    seed: int
    default_frames: int
    activation: str = "tanh"
    note: str = ""

    @property
    def n_inp(self) -> int:
        return len(self.input_ix)

    def feature_dim(self) -> int:
        d = 0
        for _, t, ix in self.features:
            if t in ("angle", "bond"):
                d += 1
            elif t == "dihedral":
                d += 1 if self.use_angle_value else 2
            else:
                d += 3 * len(ix)
        return d

    def out_dim(self) -> int:
        return self.layer_dims[-1] if self.layer_dims else self.feature_dim()

    def bytes_fwd(self) -> int:
        """Algorithmic HBM bytes per frame, forward: read x, write y."""
        return 12 * self.n_inp + 4 * self.out_dim()

    def bytes_fwd_dx(self) -> int:
        """Forward + d/dx: read x, write y, read cotangent, write dense grad_x."""
        return 24 * self.n_inp + 8 * self.out_dim()

    def mlp_flops(self) -> int:
        if not self.layer_dims:
            return 0
        return 2 * sum(a * b for a, b in zip(self.layer_dims[:-1], self.layer_dims[1:]))


def spec_c1() -> SystemSpec:
    return SystemSpec(
        name="C1", positions=ala2_positions(), input_ix=list(range(22)), align_ix=None,
        features=[("b56", "bond", [4, 5]), ("d1234", "dihedral", [0, 1, 2, 3])],
        use_angle_value=True, layer_dims=[2, 5, 3], noise=0.1, trans_sigma=0.0, rotate=False,
        seed=101, default_frames=1 << 20,
        note="alanine dipeptide, bond+dihedral FeatureLayer -> MLP [2,5,3], no alignment")


def spec_c2() -> SystemSpec:
    return SystemSpec(
        name="C2", positions=ala2_positions(), input_ix=list(range(22)), align_ix=list(ALA2_HEAVY),
        features=[("heavy", "position", list(ALA2_HEAVY))],
        use_angle_value=False, layer_dims=[30, 64, 64, 2], noise=0.3, trans_sigma=5.0, rotate=True,
        seed=202, default_frames=1 << 20,
        note="alanine dipeptide heavy-atom AlignmentLayer + position features -> MLP [30,64,64,2]")


def _chain_spec(name, n, chain_seed, seed, default_frames, hidden):
    pos = chain_positions(n, chain_seed)
    sel = list(range(0, n, 10))
    feats = [("pos", "position", sel)]
    for k in range(n // 20):
        feats.append(("d%d" % k, "dihedral", [20 * k, 20 * k + 1, 20 * k + 2, 20 * k + 3]))
    d = 3 * len(sel) + 2 * (n // 20)
    return SystemSpec(
        name=name, positions=pos, input_ix=list(range(n)), align_ix=sel, features=feats,
        use_angle_value=False, layer_dims=[d] + hidden, noise=0.2, trans_sigma=20.0, rotate=True,
        seed=seed, default_frames=default_frames,
        note="synthetic %d-atom chain, %d-atom alignment selection, positions+dihedrals" % (n, len(sel)))


def spec_c3() -> SystemSpec:
    # 2^17 frames per GPU per step (3.1 GB; SURVEY 8(d) asks for device-resident chunks of 2^17 .. 2^18 frames): the
    # persistent kernels work in 128-frame tiles per SM, so 2^15 frames were 1.7 tiles per SM and mostly pipeline fill
    return _chain_spec("C3", 2000, 2000, 303, 1 << 17, [256, 128, 2])


def spec_c5() -> SystemSpec:
    return _chain_spec("C5", 5000, 5000, 505, 1 << 16, [256, 128, 2])


def spec_small_chain(n=200, seed=77) -> SystemSpec:
    """Reduced C3-like system (fast on the CPU oracle) used by the parity tests."""
    s = _chain_spec("C3s", n, seed, 313, 512, [48, 24, 2])
    return s


SPECS = {"C1": spec_c1, "C2": spec_c2, "C3": spec_c3, "C5": spec_c5, "C3s": spec_small_chain}


def get_spec(name: str) -> SystemSpec:
    return SPECS[name]()


def default_api():
    from . import ann, feature
    return SimpleNamespace(Feature=feature.Feature, FeatureLayer=ann.FeatureLayer,
                           AlignmentLayer=ann.AlignmentLayer, PreprocessingANN=ann.PreprocessingANN,
                           MolANN=ann.MolANN, create_sequential_nn=ann.create_sequential_nn)


_ACTS = {"tanh": torch.nn.Tanh, "relu": torch.nn.ReLU, "sigmoid": torch.nn.Sigmoid}


def build_model(spec: SystemSpec, api=None, init_seed: int = 0):
    """Construct (model, universe) for ``spec`` through the molann class API."""
    api = api or default_api()
    u = Universe(spec.positions)
    input_ag = u.select_ix(spec.input_ix)
    feats = [api.Feature(nm, tp, u.select_ix(ix)) for (nm, tp, ix) in spec.features]
    flayer = api.FeatureLayer(feats, input_ag, use_angle_value=spec.use_angle_value)
    align = api.AlignmentLayer(u.select_ix(spec.align_ix), input_ag) if spec.align_ix is not None else None
    pp = api.PreprocessingANN(align, flayer)
    if spec.layer_dims is None:
        return pp, u
    torch.manual_seed(init_seed)
    net = api.create_sequential_nn(list(spec.layer_dims), _ACTS[spec.activation]())
    return api.MolANN(pp, net), u


def random_rotations(L: int, gen: torch.Generator, device, dtype=torch.float32) -> torch.Tensor:
    q = torch.randn(L, 4, generator=gen, device=device, dtype=dtype)
    q = q / q.norm(dim=1, keepdim=True)
    a, b, c, d = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c),
        2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b),
        2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d], dim=1)
    return R.reshape(L, 3, 3)


def make_frames(spec: SystemSpec, L: int, device="cpu", seed: Optional[int] = None,
                chunk: int = 1 << 16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Synthetic trajectory ``[L, n_inp, 3]`` fp32: structure + Gaussian noise (+ rigid motion)."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(spec.seed if seed is None else seed)
    base = torch.from_numpy(spec.positions[np.asarray(spec.input_ix)]).to(device)
    n = base.shape[0]
    x = out if out is not None else torch.empty(L, n, 3, device=device, dtype=torch.float32)
    for s in range(0, L, chunk):
        e = min(L, s + chunk)
        m = e - s
        fr = base.unsqueeze(0) + spec.noise * torch.randn(m, n, 3, generator=gen, device=device)
        if spec.rotate:
            R = random_rotations(m, gen, device)
            fr = torch.bmm(fr, R)
        if spec.trans_sigma > 0:
            fr = fr + spec.trans_sigma * torch.randn(m, 1, 3, generator=gen, device=device)
        x[s:e] = fr
    return x
