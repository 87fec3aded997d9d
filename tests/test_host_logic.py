"""Host-side mirror of the reference interface: masks, parameter layout, recognition, errors (CPU)."""
import re
from functools import partial

import numpy as np
import pytest
import torch

import neural_svd_b200 as N
from neural_svd_b200 import fused, operators
from conftest import build_problem, load_golden, ref_args
from oracle import nsvd_oracle as O


@pytest.mark.parametrize("L,seq,step,const", [(16, False, 1, False), (16, True, 1, False), (7, False, 3, False),
                                              (512, False, 1, True), (24, True, 1, True), (5, False, 2, True)])
def test_masks_match_reference_formula(L, seq, step, const):
    if const:
        m = N.NestedLoRAForCDK(None, L, step=step, sequential=seq, set_first_mode_const=True)
    else:
        m = N.NestedLoRA(None, L, step=step, sequential=seq)
    v, M = O.nesting_masks(L, seq, step, const)
    assert np.array_equal(m.vector_mask.numpy(), v) and np.array_equal(m.matrix_mask.numpy(), M)
    assert m.vector_mask.dtype == torch.float32 and m.name == "nestedlora"
    if not seq and step == 1 and not const:      # v_i = (L-i)/L, M_ij = min(v_i, v_j)
        assert np.allclose(v, (L - np.arange(L)) / L)
    if seq:
        assert np.array_equal(M, np.triu(np.ones_like(M)))


@pytest.mark.parametrize("name", ["hyd_small_odd", "osc_small_seq", "hyd_b128_seq_L16"])
def test_parameter_names_shapes_and_init_draws(name):
    d, cfg = load_golden(name)
    method, operator, importance, gt = build_problem(cfg, int(d["seed"]))
    names = [n for n, _ in method.named_parameters()]
    assert names == O.param_names(cfg)            # state_dict round trip with the reference
    for n, p in method.named_parameters():
        a = p.detach().numpy().astype(np.float64)
        assert np.allclose([a.sum(), (a * a).sum()], d[f"pck/{n}"], rtol=1e-12, atol=1e-12), n
    assert not dict(method.named_parameters())["model.base.feature_map._B"].requires_grad
    assert np.allclose(gt, d["gt"][:cfg.neigs])


def test_describe_operator_and_importance():
    d, cfg = load_golden("osc_small_seq")
    method, operator, importance, _ = build_problem(cfg, 0)
    od = operators.describe_operator(operator)
    assert od == dict(potential=1, pot_coef=1.0, pot_coef2=0.0, scale_kinetic=1.0, op_scale=1.0, op_shift=16.0,
                      fd_eps=0.0)
    # laplacian_eps > 0 selects the finite-difference Laplacian (diff_ops.py:7), on the mirror and on a duck-typed
    # reference-style object that only keeps it inside its VectorizedLaplacian
    _, op_fd, _, _ = build_problem(cfg, 0, laplacian_eps=0.01)
    assert operators.describe_operator(op_fd)["fd_eps"] == pytest.approx(0.01)
    import types
    ref_like = types.SimpleNamespace(local_potential_ftn=op_fd.operator.local_potential_ftn, scale_kinetic=1.0,
                                     n_particles=1, laplacian=types.SimpleNamespace(eps=0.1))
    assert operators.describe_operator(operators.OperatorWrapper(ref_like, 1.0, 0.0))["fd_eps"] == pytest.approx(0.1)
    assert operators.describe_importance(importance) == dict(importance=0, sigma=4.0)
    md = fused.describe_model(method)
    assert md["L"] == 4 and md["Mff"] == 64 and md["scales"] is not None and md["hard_mul_const"] == 0.5
    # the reference's importance is a closure over a MultivariateNormal (main_pde.py:94-100)
    from torch.distributions import MultivariateNormal
    mvn = MultivariateNormal(loc=torch.zeros(2), covariance_matrix=16.0 ** 2 * torch.eye(2))
    closure = lambda x: mvn.log_prob(x.view(x.shape[0], -1)).exp().view(-1, 1)  # noqa: E731
    assert operators.describe_importance(closure) == dict(importance=0, sigma=16.0)
    x = 16 * torch.randn(8, 2)
    assert torch.allclose(N.GaussianImportance(16.0)(x), closure(x), rtol=1e-5)
    # importance=None is the reference's un-weighted operator call (diff_ops.py:10-11)
    assert operators.describe_importance(None)["importance"] == 3


def test_other_samplers_potentials_and_masks_are_recognised():
    """main_pde.py:101-118 closures read `args`; problems.py:30-73 potentials; boundary.py:16-37."""
    from types import SimpleNamespace
    from torch.distributions import Laplace
    args = SimpleNamespace(sampling_mode="laplacian", sampling_scale=3.0, ndim=2, n_particles=1)

    def importance_train(x):
        return Laplace(torch.zeros(2), args.sampling_scale * torch.ones(2)).log_prob(x).sum(-1).exp().view(-1, 1)
    assert operators.describe_importance(importance_train) == dict(importance=1, sigma=3.0)
    x = 3 * torch.randn(8, 2)
    assert torch.allclose(N.LaplaceImportance(3.0)(x), importance_train(x), rtol=1e-5)
    args.sampling_mode = "uniform"
    assert operators.describe_importance(importance_train) == dict(importance=2, sigma=3.0)
    assert torch.allclose(N.UniformImportance(3.0)(x), torch.full((8, 1), 1 / 36.0))

    d, cfg = load_golden("molion_laplace_boxexp_mask")
    method, operator, importance, gt = build_problem(cfg, int(d["seed"]))
    assert gt is None
    od = operators.describe_operator(operator)
    assert (od["potential"], od["pot_coef"], od["pot_coef2"]) == (2, 2.0, 1.5)
    md = fused.describe_model(method)
    assert md["scales"] is not None and (md["box_mode"], md["box_lim"]) == (2, 12.0)
    d, cfg = load_golden("cosine_uniform_detff")
    method, operator, importance, gt = build_problem(cfg, int(d["seed"]))
    assert np.allclose(gt, d["gt"]) and fused.describe_model(method)["Mff"] == 64
    assert operators.describe_operator(operator)["potential"] == 4
    d, cfg = load_golden("well_uniform_boxsqrt")
    method, operator, importance, gt = build_problem(cfg, int(d["seed"]))
    assert np.allclose(gt, d["gt"]) and fused.describe_model(method)["box_mode"] == 1
    xx = torch.tensor([[0.3, -1.2]])
    assert torch.allclose(N.hydrogen_mol_ion_potential(xx, R=1.5, charge=2.0),
                          N.hydrogen_potential(xx - torch.tensor([0.0, 1.5]), 2.0)
                          + N.hydrogen_potential(xx + torch.tensor([0.0, 1.5]), 2.0))


def test_unsupported_configurations_raise():
    with pytest.raises(NotImplementedError):
        operators.describe_operator(N.OperatorWrapper(N.NegativeHamiltonian(lambda x: x), 1.0, 0.0))
    with pytest.raises(NotImplementedError):
        operators.describe_importance(lambda x: x)
    a = ref_args(O.PathConfig.hydrogen(neigs=2, fourier_mapping_size=8))
    a.parallel = False
    with pytest.raises(NotImplementedError):
        N.get_wavefunctions(a)
    a.parallel, a.nonlinearity = True, "relu"
    with pytest.raises(NotImplementedError):
        N.get_wavefunctions(a)
    with pytest.raises(NotImplementedError):
        N.NestedLoRA(None, 4).compute_loss_operator(None, None, evd=False)
    # ndim = 3 exists for the potentials the reference runs there (problems.py:62-71); its oscillator / well / cosine
    # branches assert other dimensions, and so does the mirror; other dimensions and several particles raise
    a = ref_args(O.PathConfig.oscillator(neigs=2, fourier_mapping_size=8, ndim=3))
    with pytest.raises(NotImplementedError):
        N.get_problem(a)
    a = ref_args(O.PathConfig.hydrogen(neigs=2, fourier_mapping_size=8, ndim=4))
    with pytest.raises(NotImplementedError):
        N.get_problem(a)


def test_ndim3_problem_is_recognised():
    from neural_svd_b200 import fused
    d, cfg = load_golden("hyd3d_small")
    method, operator, importance, gt = build_problem(cfg, int(d["seed"]))
    md = fused.describe_model(method)
    assert md["ndim"] == 3 and tuple(md["Bff"].shape) == (3, cfg.fourier_mapping_size)
    assert np.allclose(gt, d["gt"])                                  # Hydrogen3D.get_eigvals, scaled (problems.py:126-128)
    assert len(operators.hydrogen3d_eigvals(16)) == 14               # the reference's enumeration stops at n = 3
    assert operators.describe_importance(importance, 3) == dict(importance=0, sigma=cfg.sampling_scale)
    with pytest.raises(NotImplementedError):
        operators.describe_importance(N.GaussianImportance(1.0, 2), 3)
    with pytest.raises(NotImplementedError, match=r"\(B, 3\)"):
        fused._prep_x(torch.zeros(4, 2), torch.device("cpu"), 3)


def test_no_cpu_fallback():
    d, cfg = load_golden("hyd_small_odd")
    method, operator, importance, _ = build_problem(cfg, 0)
    x = torch.from_numpy(d["x"])
    with pytest.raises(RuntimeError, match="no CPU path"):
        method.compute_loss_operator(operator, x, importance=importance)
    with pytest.raises(RuntimeError, match="no CPU path"):
        N.NestedLoRAForCDK(None, 4).compute_loss(torch.randn(8, 4), torch.randn(8, 4))


def test_product_never_imports_oracle():
    import os
    root = os.path.dirname(os.path.abspath(N.__file__))
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_shard_points_tiles_global_halves():
    x = torch.arange(16.0).view(8, 2)
    parts = [N.shard_points(x, r, 2) for r in range(2)]
    first = torch.cat([p[:2] for p in parts])
    second = torch.cat([p[2:] for p in parts])
    assert torch.equal(first, x[:4]) and torch.equal(second, x[4:])


def test_potentials_match_reference_formulas():
    x = torch.tensor([[3.0, 4.0]])
    assert torch.allclose(N.hydrogen_potential(x, charge=2.0), torch.tensor([[-0.4]]))
    assert torch.allclose(N.harmonic_oscillator_potential(x, k=0.5), torch.tensor([[12.5]]))


def test_sketchy_encoder_shapes_and_cpu_guards():
    """the mirror of main_sketchy.py:107-115 has the script's shapes; device-only helpers refuse CPU."""
    enc = N.get_sketchy_encoder()
    shapes = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    assert shapes == {"backbones.x.0.weight": (8192, 512), "backbones.x.0.bias": (8192,),
                      "backbones.x.2.weight": (512, 8192), "backbones.x.2.bias": (512,),
                      "backbones.y.0.weight": (8192, 512), "backbones.y.0.bias": (8192,),
                      "backbones.y.2.weight": (512, 8192), "backbones.y.2.bias": (512,)}
    assert enc.output_dims == {"x": 512, "y": 512} and enc.mu == 16.0
    z = 10 * torch.randn(32, 512)
    out = N.normalize(z, 4.0, "l2_ball")
    assert float(out.norm(dim=1).max()) <= 4.0 + 1e-4            # rows outside the ball are projected onto it
    small = 0.01 * torch.randn(4, 512)
    assert torch.equal(N.normalize(small, 4.0, "l2_ball"), small)   # rows inside are untouched
    with pytest.raises(RuntimeError, match="no CPU path"):
        N.sample_points(16, "uniform", 1.0, seed=0, device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        N.sample_gaussian(16, 1.0, seed=0, device="cpu")
    from torch.distributions import Laplace
    lap = Laplace(torch.zeros(2), 2.5 * torch.ones(2))
    closure = lambda x: lap.log_prob(x).sum(-1).exp().view(-1, 1)  # noqa: E731
    assert operators.describe_importance(closure) == dict(importance=1, sigma=2.5)


def test_description_and_mask_caches_follow_their_sources():
    """The per-step host work (structural checks of the module, the reference's CPU-resident masks moved to the device,
    nestedlora.py:87-88) is cached; the caches must notice every change of what they describe."""
    cfg = O.PathConfig.hydrogen()
    method, operator, importance, _ = build_problem(cfg, 0)
    md = fused.describe_model(method)
    assert fused.describe_model(method) is md                      # same module, same parameter objects: a hit
    assert md["ws"][0] is method.model.base.ws[0]                  # the caller's own Parameter objects, never copies
    method.model.hard_mul_const = 3.0                              # a scalar setting changes
    md2 = fused.describe_model(method)
    assert md2 is not md and md2["hard_mul_const"] == 3.0
    method.model.base.ws[1] = torch.nn.Parameter(torch.zeros_like(method.model.base.ws[1]))   # a parameter is replaced
    md3 = fused.describe_model(method)
    assert md3 is not md2 and md3["ws"][1] is method.model.base.ws[1]
    method.model.base.double()                                     # dtype changes under the same Parameter objects
    with pytest.raises(NotImplementedError):
        fused.describe_model(method)

    method, operator, importance, _ = build_problem(cfg, 0)
    dev = torch.device("cpu")
    v, M = fused._nesting_masks(method, dev)
    assert fused._nesting_masks(method, dev)[0] is v
    assert v is not method.vector_mask and torch.equal(v, method.vector_mask)     # never an alias of the caller's tensor
    method.vector_mask.mul_(2.0)                                   # in-place edit: the version counter moves
    v2, _ = fused._nesting_masks(method, dev)
    assert v2 is not v and torch.equal(v2, method.vector_mask)
    method.matrix_mask = method.matrix_mask.clone() + 1.0          # replaced object
    _, M2 = fused._nesting_masks(method, dev)
    assert torch.equal(M2, method.matrix_mask)
    method.register_eigvals(torch.arange(cfg.neigs, dtype=torch.float32))         # sort_indices appear (nestedlora.py:202-208)
    v3, M3 = fused._nesting_masks(method, dev)
    si = method.sort_indices
    assert torch.equal(v3[si], method.vector_mask) and torch.equal(M3[si][:, si], method.matrix_mask)
    method.reset_eigvals()
    assert torch.equal(fused._nesting_masks(method, dev)[0], method.vector_mask)

    from torch.distributions import MultivariateNormal
    mvn = MultivariateNormal(torch.zeros(2), 16.0 * torch.eye(2))

    def importance_train(x):
        return mvn.log_prob(x).exp().view(-1, 1)
    d = operators.describe_importance(importance_train)
    assert d == dict(importance=0, sigma=4.0) and operators.describe_importance(importance_train) == d
    d["sigma"] = -1.0                                              # the caller's copy is not the cached one
    assert operators.describe_importance(importance_train)["sigma"] == 4.0
