"""Parity of the sm_100a path (through the Python API -> C-ABI) against the golden vectors of the
UNMODIFIED reference and against the CPU oracle.  Tolerance (north star): loss and parameter
gradients within 1e-4 relative (Frobenius, per tensor)."""
import numpy as np
import pytest
import torch

import neural_svd_b200 as N
from conftest import build_problem, golden_grad_errors, load_golden, rel
from oracle import nsvd_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-4
ENGINES = ["fp32", "f16x3"]
PDE_CASES = ["hyd_small_odd", "osc_small_seq", "hyd_small_sorted", "hyd_b128_seq_L16", "osc_b512_jnt_L16", "hyd_b512_jnt_L16",
             "hyd_b64_jnt_L64",
             # infinite well / cosine / H2+ potentials, uniform / Laplace / no importance,
             # Dirichlet box masks (sqrt, exp; alone and under the exp mask), deterministic Fourier features
             "well_uniform_boxsqrt", "cosine_uniform_detff", "molion_laplace_boxexp_mask", "osc_no_importance",
             # ndim = 3 (five streams; both engine settings resolve to the fp32 CUDA-core engine, see test_ndim3_*)
             "hyd3d_small", "molion3d_laplace_boxexp"]


def _step(name, engine):
    d, cfg = load_golden(name)
    N.set_engine(engine)
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda")
    if "eigvals" in d:                       # nestedlora.py:202-206: output columns sorted by the registered eigenvalues
        method.register_eigvals(d["eigvals"])
    x = torch.from_numpy(d["x"]).cuda()
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    grads = {n: p.grad.detach().cpu().numpy() for n, p in method.named_parameters() if p.grad is not None}
    return d, cfg, method, float(loss.detach()), aux, grads


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", PDE_CASES)
def test_step_matches_reference_golden(name, engine):
    d, cfg, method, loss, aux, grads = _step(name, engine)
    assert abs(loss - float(d["loss64"])) <= TOL * abs(float(d["loss64"]))
    assert rel(aux["f"].cpu().numpy(), d["f64"]) < TOL
    assert rel(aux["Tf"].cpu().numpy(), d["Tf64"]) < TOL
    assert aux["eigvals"] is None and aux["f"].shape == d["f64"].shape
    names = [n for n in O.param_names(cfg) if n != "model.base.feature_map._B"]
    assert sorted(grads) == sorted(names)                      # _B gets no gradient
    errs = golden_grad_errors(d, names, grads)
    assert len(errs) == len(names)
    assert max(errs.values()) < TOL, errs


PDE_FD = ["hyd_small_fd0p1", "osc_small_fd0p01", "hyd_b512_jnt_L16_fd0p01", "hyd3d_small_fd0p05"]


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", PDE_FD)
def test_finite_difference_step_matches_reference(name, engine):
    # laplacian_eps > 0 (every shipped script: hydrogen.sh:20): VectorizedLaplacian.approx_laplacian, diff_ops.py:25-52.
    # Parity is against the reference's FD path in fp64.  Second differences divide round-off by eps^2, so in fp32 the
    # REFERENCE itself is this far from its own fp64 result (stored in the fixture): Tf 3e-4 (small cases), 4e-2
    # (hydrogen L=16, eps = 0.01).  The bar: within 3x that self-noise for the fp32 engine, 10x for the tensor-core
    # engine (22-bit operands), and never worse than 1e-4 where the reference's noise is below it.
    d, cfg = load_golden(name)
    eps = float(d["laplacian_eps"])
    N.set_engine(engine)
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda", laplacian_eps=eps)
    x = torch.from_numpy(d["x"]).cuda()
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    grads = {n: p.grad.detach().cpu().numpy() for n, p in method.named_parameters() if p.grad is not None}
    k = 3.0 if engine == "fp32" else 10.0
    e_loss = abs(float(loss.detach()) / float(d["loss64"]) - 1)
    e_f, e_tf = rel(aux["f"].cpu().numpy(), d["f64"]), rel(aux["Tf"].cpu().numpy(), d["Tf64"])
    errs = golden_grad_errors(d, list(grads), grads)
    print(f"[{engine}] {name}: loss {e_loss:.1e} (ref self {float(d['loss_self']):.1e}) f {e_f:.1e} Tf {e_tf:.1e} "
          f"(ref self {float(d['tf_self']):.1e}) grads max {max(errs.values()):.1e}")
    assert e_f < TOL                                      # f is the central evaluation: no cancellation
    assert e_tf < max(TOL, k * float(d["tf_self"]))
    # the loss is ONE number: the reference's own deviation is a single noise draw, so its bar is 10x for both engines
    assert e_loss < max(TOL, 10.0 * float(d["loss_self"]))
    for n, e in errs.items():
        assert e < max(TOL, k * float(d[f"gself/{n}"])), (n, e, float(d[f"gself/{n}"]))
    # the same model with laplacian_eps = 0 gives the exact-Laplacian result: the two modes really differ
    method0, operator0, _, _ = build_problem(cfg, int(d["seed"]), "cuda")
    _, aux0 = method0.compute_loss_operator(operator0, x, importance=importance)
    assert rel(aux0["Tf"].cpu().numpy(), aux["Tf"].cpu().numpy()) > 1e-7


def test_ndim3_runs_on_the_cuda_core_engine_across_micro_batches():
    """ndim = 3 (problems.py:62-71): fresh seeded inputs against the oracle at a size that spans two micro-batches of the
    fp32 engine with a ragged tail, exp mask + hard_mul_const, perturbed biases; the tensor-core setting resolves to the
    same CUDA-core engine (and says so once), the graph-captured step and the 3D device sampler work."""
    import warnings
    from neural_svd_b200 import fused
    cfg = O.PathConfig(potential="hydrogen_mol_ion", ndim=3, neigs=5, fourier_mapping_size=48, fourier_scale=0.3,
                       operator_scale=10.0, sampling_scale=3.0, apply_exp_mask=True, exp_mask_init_scale=6.0,
                       hard_mul_const=0.7, hydrogen_mol_ion_R=0.8, sequential=True)
    N.set_engine("f16x3")
    method, operator, importance, _ = build_problem(cfg, 31, "cuda")
    g = torch.Generator().manual_seed(6)
    x = cfg.sampling_scale * torch.randn(2048 + 173, 3, generator=g)
    with torch.no_grad():
        for b in method.model.base.bs:
            b.add_(0.1 * torch.randn(b.shape, generator=g).to(b.device))
    params = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in method.named_parameters()}
    r = O.train_step(x.numpy().astype(np.float64), params, cfg)
    fused._warned_3d = False
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        loss, aux = method.compute_loss_operator(operator, x.cuda(), importance=importance)
    assert any("fp32 CUDA-core engine" in str(m.message) for m in w)
    assert fused.engine_for(fused.describe_model(method)) == 0
    loss.backward()
    assert abs(float(loss.detach()) - r["loss"]) <= TOL * abs(r["loss"])
    assert rel(aux["f"].cpu().numpy(), r["f"]) < TOL and rel(aux["Tf"].cpu().numpy(), r["Tf"]) < TOL
    grads = {n: p.grad.clone() for n, p in method.named_parameters() if p.grad is not None}
    for n, gr in grads.items():
        assert rel(gr.cpu().numpy(), r["grads"][n]) < TOL, n
    step = N.GraphedOperatorStep(method, operator, importance, x.shape[0])
    l2 = step(x.cuda())
    assert abs(float(l2) - float(loss.detach())) < 1e-5 * abs(float(loss.detach()))
    for n, p in method.named_parameters():
        if n in grads:
            assert rel(p.grad.cpu().numpy(), grads[n].cpu().numpy()) < 1e-5, n
    xs = N.sample_points(4001, "gaussian", 2.0, seed=3, ndim=3)
    assert xs.shape == (4001, 3) and abs(float(xs.std()) - 2.0) < 0.05 and abs(float(xs.mean())) < 0.05
    assert torch.equal(xs, N.sample_gaussian(4001, 2.0, seed=3, ndim=3))


@pytest.mark.parametrize("neigs", [6, 5])
@pytest.mark.parametrize("engine", ENGINES)
def test_step_matches_oracle_on_fresh_inputs(engine, neigs):
    # seeded inputs that are in no fixture, at a ragged size (B not a multiple of the 128-row tile); an odd number
    # of copies exercises the zero-filled second copy of the last CTA pair in the weight-gradient GEMM
    cfg = O.PathConfig.oscillator(neigs=neigs, fourier_mapping_size=40, sequential=False, step=2)
    N.set_engine(engine)
    method, operator, importance, _ = build_problem(cfg, 77, "cuda")
    g = torch.Generator().manual_seed(5)
    x = cfg.sampling_scale * torch.randn(301, 2, generator=g)
    # perturb biases / scales so that they are exercised (reference initialises biases to zero)
    with torch.no_grad():
        for i, b in enumerate(method.model.base.bs):
            b.add_(0.1 * torch.randn(b.shape, generator=g).to(b.device))
        method.model.boundary_mask.scales.mul_(1 + 0.2 * torch.rand(neigs, generator=g).to("cuda"))
    params = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in method.named_parameters()}
    r = O.train_step(x.numpy().astype(np.float64), params, cfg)
    loss, aux = method.compute_loss_operator(operator, x.cuda(), importance=importance)
    loss.backward()
    assert abs(float(loss.detach()) - r["loss"]) <= TOL * abs(r["loss"])
    assert rel(aux["Tf"].cpu().numpy(), r["Tf"]) < TOL
    for n, p in method.named_parameters():
        if p.grad is not None:
            assert rel(p.grad.cpu().numpy(), r["grads"][n]) < TOL, n


def test_grad_output_scaling_and_accumulation():
    d, cfg = load_golden("hyd_small_odd")
    N.set_engine("fp32")
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda")
    x = torch.from_numpy(d["x"]).cuda()
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    g1 = method.model.base.ws[1].grad.clone()
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    (3.0 * loss).backward()                                  # accumulates 3x on top of 1x
    assert rel(method.model.base.ws[1].grad.cpu().numpy(), 4.0 * g1.cpu().numpy()) < 1e-5


def test_operator_protocol_and_model_forward():
    d, cfg = load_golden("osc_small_seq")
    N.set_engine("fp32")
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda")
    x = torch.from_numpy(d["x"]).cuda()
    Tf, f = operator(method, x, importance=importance)       # examples/__init__.py:7-9 protocol
    assert rel(Tf.cpu().numpy(), d["Tf64"]) < TOL and rel(f.cpu().numpy(), d["f64"]) < TOL
    vals = method(x)                                          # NestedLoRA.forward -> WaveFunctions.forward
    params = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in method.named_parameters()}
    u = O.forward_streams(d["x"].astype(np.float64), params, cfg)
    r = np.sqrt((d["x"].astype(np.float64) ** 2).sum(1))[:, None]
    want = cfg.hard_mul_const * u[0] * np.exp(-r / params["model.boundary_mask.scales"][None, :])
    assert rel(vals.cpu().numpy(), want) < TOL


@pytest.mark.parametrize("seq", [False, True])
@pytest.mark.parametrize("B,L", [(97, 5), (1000, 16), (513, 33), (2048, 64)])
def test_standalone_loss_function_matches_oracle(B, L, seq):
    # NestedLoRALossFunctionEVD.apply(f, Tf, f1, f2, v, M): K2 + K3 on random inputs (nestedlora.py:67-111)
    g = torch.Generator().manual_seed(B + L)
    f = torch.randn(B, L, generator=g)
    Tf = torch.randn(B, L, generator=g)
    v, M = O.nesting_masks(L, seq)
    loss_o, lam1, lam2 = O.loss_forward(f.numpy().astype(np.float64), Tf.numpy().astype(np.float64), v, M)
    dF_o = O.loss_dF(f.numpy().astype(np.float64), Tf.numpy().astype(np.float64), v, M, lam1, lam2)
    fc = f.cuda().requires_grad_()
    f1, f2 = torch.chunk(fc, 2)
    loss = N.NestedLoRALossFunctionEVD.apply(fc, Tf.cuda(), f1, f2, torch.from_numpy(v), torch.from_numpy(M))
    loss.backward()
    assert abs(float(loss.detach()) - loss_o) < 1e-5 * abs(loss_o)
    assert rel(fc.grad.cpu().numpy(), dF_o) < 1e-5


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("name", ["cdk_small_seq", "cdk_small_noconst"])
def test_cdk_matches_reference_golden(name, engine):
    d, _ = load_golden(name)
    N.set_engine(engine)
    m = N.NestedLoRAForCDK(None, int(d["L"]), step=1, sequential=bool(d["sequential"]),
                           set_first_mode_const=bool(d["const"]))
    f = torch.from_numpy(d["f"]).cuda().requires_grad_()
    g = torch.from_numpy(d["g"]).cuda().requires_grad_()
    loss, lop, lmet, rsj, rsi = m.compute_loss(f, g)
    loss.backward()
    assert abs(float(loss.detach()) - float(d["loss64"])) < TOL * abs(float(d["loss64"]))
    assert abs(float(lop) - float(d["lop64"])) < TOL * abs(float(d["lop64"]))
    assert abs(float(lmet) - float(d["lmet64"])) < TOL * abs(float(d["lmet64"]))
    assert rel(f.grad.cpu().numpy(), d["gf64"]) < TOL and rel(g.grad.cpu().numpy(), d["gg64"]) < TOL
    assert rel(rsj.cpu().numpy(), d["rsj64"]) < TOL and rel(rsi.cpu().numpy(), d["rsi64"]) < TOL
    assert rsi.shape == (f.shape[0] ** 2 - f.shape[0],)


@pytest.mark.parametrize("engine", ENGINES)
def test_cdk_full_size_config5(engine):
    d, _ = load_golden("cdk_b4096_L512")
    N.set_engine(engine)
    g = torch.Generator().manual_seed(int(d["seed"]))
    f = torch.randn(int(d["B"]), int(d["L"]), generator=g).cuda().requires_grad_()
    gg = torch.randn(int(d["B"]), int(d["L"]), generator=g).cuda().requires_grad_()
    m = N.NestedLoRAForCDK(None, int(d["L"]))
    loss, lop, lmet, rsj, rsi = m.compute_loss(f, gg)
    loss.backward()
    assert abs(float(loss.detach()) - float(d["loss64"])) < TOL * abs(float(d["loss64"]))
    assert rel(f.grad.cpu().numpy().reshape(-1)[d["gidx"]], d["gfval"]) < TOL
    assert rel(gg.grad.cpu().numpy().reshape(-1)[d["gidx"]], d["ggval"]) < TOL
    assert rel(rsj.cpu().numpy(), d["rsj64"]) < TOL
    assert rel(rsi.cpu().numpy()[d["rsi_idx"]], d["rsi_val"]) < TOL


def test_cdk_cabi_planes_ready_and_finalize_scratch():
    """ABI 5 contracts of the CDK entry points, called directly (include/nsvd.h): bwd / offdiag with planes_ready = 0
    rebuild the operand planes from f, g (a scribbled work buffer must not matter) and give what planes_ready = 1 gives on
    the untouched buffer of the forward; finalize only needs its scratch to be writable (contents irrelevant) and is
    deterministic; all against the golden vectors of the reference (nestedlora.py:270-332)."""
    import ctypes as C
    from neural_svd_b200 import _lib
    d, _ = load_golden("cdk_small_seq")
    lib = _lib.load()
    eng = _lib.ENGINES["f16x3"]
    f, g = torch.from_numpy(d["f"]).cuda(), torch.from_numpy(d["g"]).cuda()
    B, L = f.shape
    fc = int(bool(d["const"]))
    Lp = L + fc
    m = N.NestedLoRAForCDK(None, L, step=1, sequential=bool(d["sequential"]), set_first_mode_const=bool(fc))
    v, Mm = m.vector_mask.cuda(), m.matrix_mask.cuda().contiguous()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nwork = lib.nsvd_cdk_work_bytes(B, L, fc, eng)
    work = torch.empty(nwork, dtype=torch.uint8, device="cuda")
    terms = torch.empty(2 * Lp * Lp + 1, device="cuda")
    rsj = torch.empty(B, device="cuda")
    _lib.check(lib.nsvd_cdk_fwd(_lib.ptr(f), _lib.ptr(g), _lib.ptr(v), B, L, fc, eng, _lib.ptr(terms), _lib.ptr(rsj),
                                _lib.ptr(work), nwork, st), "fwd")
    outs = []
    for fill in (0x00, 0xFF):                        # the scratch contents do not matter
        scratch = torch.full((_lib.CDK_FINALIZE_SCRATCH,), fill, dtype=torch.uint8, device="cuda")
        losses, coef = torch.empty(3, device="cuda"), torch.empty(2 * Lp * Lp, device="cuda")
        _lib.check(lib.nsvd_cdk_finalize(_lib.ptr(terms), _lib.ptr(Mm), Lp, B, _lib.ptr(losses), _lib.ptr(coef),
                                         _lib.ptr(scratch), st), "finalize")
        outs.append((losses.clone(), coef.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    losses, coef = outs[0]
    assert abs(float(losses[0]) - float(d["loss64"])) < TOL * abs(float(d["loss64"]))
    one = torch.ones((), device="cuda")

    def bwd_and_offdiag(ready):
        gf, gg = torch.empty_like(f), torch.empty_like(g)
        rsi = torch.empty(B * B - B, device="cuda")
        _lib.check(lib.nsvd_cdk_offdiag(_lib.ptr(f), _lib.ptr(g), B, L, fc, eng, _lib.ptr(rsi), _lib.ptr(work), nwork,
                                        ready, st), "offdiag")
        _lib.check(lib.nsvd_cdk_bwd(_lib.ptr(f), _lib.ptr(g), _lib.ptr(v), _lib.ptr(coef), _lib.ptr(one), B, L, fc, B, eng,
                                    _lib.ptr(gf), _lib.ptr(gg), _lib.ptr(work), nwork, ready, st), "bwd")
        return gf, gg, rsi
    a = bwd_and_offdiag(1)                           # the planes of the forward
    work.fill_(0x7F)                                 # scribble: planes_ready = 0 must rebuild them
    b = bwd_and_offdiag(0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert rel(a[0].cpu().numpy(), d["gf64"]) < TOL and rel(a[1].cpu().numpy(), d["gg64"]) < TOL
    assert rel(a[2].cpu().numpy(), d["rsi64"]) < TOL
    with pytest.raises(RuntimeError):                # a misaligned scratch pointer is refused, not dereferenced
        bad = C.c_void_p(scratch.data_ptr() + 4)
        _lib.check(lib.nsvd_cdk_finalize(_lib.ptr(terms), _lib.ptr(Mm), Lp, B, _lib.ptr(losses), _lib.ptr(coef), bad, st),
                   "finalize")


@pytest.mark.parametrize("B,L,b1", [(5000, 16, 2500), (4096 + 77, 16, 2001), (3000, 64, 1500), (1001, 40, 333), (777, 24, 400),
                                     (513, 33, 257), (70000, 64, 35000), (300000, 16, 150001)])
def test_k2_k3_kernels_against_fp64(B, L, b1):
    """K2 (nsvd_gram_reduce) and K3 (nsvd_loss_dF) through the C-ABI on every kernel variant behind them - the L = 16
    one-launch Gram and the pipelined / block-staged dF kernels, the mma.sync kernels for 16 < L <= 64 (L a multiple of 4,
    padded tiles, ragged chunks, halves of odd size) and the CUDA-core fallbacks - against an fp64 evaluation of
    nestedlora.py:57-64, 98-111.  Both arithmetic engines share these kernels, so engine-vs-engine tests cannot see them."""
    import ctypes as C
    from neural_svd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(B + L)
    F = torch.randn(B, L, generator=g).cuda()
    TF = (3.0 * torch.randn(B, L, generator=g)).cuda()
    v = torch.rand(L, generator=g).cuda() + 0.1
    coef = torch.randn(2 * L * L + 1, generator=g).cuda()
    gs = torch.tensor(0.7).cuda()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    terms = torch.empty(2 * L * L + 5).cuda()
    part = torch.empty(lib.nsvd_gram_partials_bytes(B, L), dtype=torch.uint8, device="cuda")
    _lib.check(lib.nsvd_gram_reduce(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), B, L, b1, _lib.ptr(terms), _lib.ptr(part), st),
               "nsvd_gram_reduce")
    dF = torch.empty_like(F)
    _lib.check(lib.nsvd_loss_dF(_lib.ptr(F), _lib.ptr(TF), _lib.ptr(v), _lib.ptr(coef), _lib.ptr(gs), B, L, b1, B,
                                _lib.ptr(dF), st), "nsvd_loss_dF")
    Fd, Td, vd, cd = F.double(), TF.double(), v.double(), coef.double()
    G1, G2 = Fd[:b1].T @ Fd[:b1], Fd[b1:].T @ Fd[b1:]
    ops = (vd * Fd * Td).sum()
    t = terms.double()
    LL = L * L
    assert rel(t[:LL].cpu().numpy(), G1.reshape(-1).cpu().numpy()) < 2e-6
    assert rel(t[LL:2 * LL].cpu().numpy(), G2.reshape(-1).cpu().numpy()) < 2e-6
    assert abs(float(t[2 * LL]) - float(ops)) < 2e-6 * float((vd * Fd * Td).abs().sum())
    ref = -(4.0 / B) * vd * Td
    ref[:b1] += Fd[:b1] @ cd[:LL].reshape(L, L)
    ref[b1:] += Fd[b1:] @ cd[LL:2 * LL].reshape(L, L)
    ref *= 0.7
    assert rel(dF.double().cpu().numpy(), ref.cpu().numpy()) < 2e-6


def test_large_batch_properties():
    # BASELINE-size inputs where the oracle is too slow: size-independent properties instead.
    # (1) permuting points inside each half leaves loss and gradients unchanged;
    # (2) data-parallel split identity: terms of the two halves of a batch add up to the whole.
    cfg = O.PathConfig.hydrogen()
    N.set_engine("fp32")
    method, operator, importance, _ = build_problem(cfg, 0, "cuda")
    g = torch.Generator().manual_seed(9)
    B = 4096
    x = (cfg.sampling_scale * torch.randn(B, 2, generator=g)).cuda()
    loss, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward()
    g0 = method.model.base.ws[0].grad.clone()
    method.zero_grad()
    perm = torch.cat([torch.randperm(B // 2, generator=g), B // 2 + torch.randperm(B // 2, generator=g)]).cuda()
    loss2, aux2 = method.compute_loss_operator(operator, x[perm], importance=importance)
    loss2.backward()
    assert abs(float(loss.detach()) - float(loss2)) < 1e-5 * abs(float(loss.detach()))
    assert torch.allclose(aux2["f"], aux["f"][perm], rtol=1e-5, atol=1e-6)
    assert rel(method.model.base.ws[0].grad.cpu().numpy(), g0.cpu().numpy()) < 1e-5


def test_spectrum_evd_matches_oracle():
    # methods/spectrum.py:29-102 on a small validation grid that contains the origin (hydrogen: V = -inf there)
    cfg = O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64)
    N.set_engine("fp32")
    method, operator, importance, _ = build_problem(cfg, 11, "cuda")
    lim, eps = 4.0, 0.25
    ax = np.arange(-lim, lim, eps)
    xx, yy = np.meshgrid(ax, ax)
    grid = torch.tensor(np.stack([xx.ravel(), yy.ravel()], 1)).float()
    assert (grid.abs().sum(1) == 0).any()

    def loader():
        for i in range(0, len(grid), 300):
            yield grid[i:i + 300], 0.0

    def importance_val(x):
        return (1 / (2 * lim) ** 2 * torch.ones(x.shape[0], 1)).to(x.device).float()

    out = N.compute_spectrum_evd(method, loader(), operator, importance_train=importance,
                                 importance_val=importance_val, device="cuda")
    params = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in method.named_parameters()}
    ref = O.spectrum_evd(grid.numpy().astype(np.float64), params, cfg, lim, chunk=300)
    assert rel(out["cov"], ref["cov"]) < TOL and rel(out["quad"], ref["quad"]) < TOL
    assert rel(out["eigvals"], ref["eigvals"]) < TOL and rel(out["norms"], ref["norms"]) < TOL
    assert out["eigfuncs"].shape == (len(grid), 4)


@pytest.mark.parametrize("engine", ["fp32", "f16x3"])
@pytest.mark.parametrize("name", ["spec_hyd_small", "spec_osc_small"])
def test_spectrum_evd_matches_reference_fixture(name, engine):
    # methods/spectrum.py:29-102 as run by the UNMODIFIED reference (oracle/make_golden.py SPECTRUM_CASES): plain,
    # normalize, and normalize + sort + post_align outputs on a validation grid that contains the origin
    from test_oracle_golden import _spec_grid, spec_close
    d, cfg = load_golden(name)
    N.set_engine(engine)
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda")
    grid, lim, chunk = torch.from_numpy(_spec_grid(d)), float(d["lim"]), int(d["chunk"])

    def loader():
        for i in range(0, len(grid), chunk):
            yield grid[i:i + chunk], 0.0

    def importance_val(x):
        return (1 / (2 * lim) ** 2 * torch.ones(x.shape[0], 1)).to(x.device).float()

    for flags, kw in dict(plain={}, norm=dict(normalize=True),
                          all=dict(normalize=True, sort=True, post_align=True)).items():
        out = N.compute_spectrum_evd(method, loader(), operator, importance_train=importance,
                                     importance_val=importance_val, device="cuda", **kw)
        # the aligned outputs go through two eigendecompositions of 4x4 matrices with condition ~1e3
        spec_close(out, d, flags + "64", TOL, aligned_tol=20 * TOL)
    N.set_engine("f16x3")


def test_fused_rmsprop_ema_matches_torch():
    # examples/utils.py:48-57 + CosineAnnealingLR + torch_ema semantics (operator/__init__.py:34-36,69-73)
    g = torch.Generator().manual_seed(0)
    shapes = [(16, 128, 64), (16, 128, 1), (16,), (3, 5, 7)]
    p_ref = [torch.randn(s, generator=g).cuda().requires_grad_() for s in shapes]
    p_our = [p.detach().clone().requires_grad_() for p in p_ref]
    S, decay = 25, 0.995
    opt = torch.optim.RMSprop(p_ref, lr=1e-3, alpha=0.999, eps=1e-10, weight_decay=0, momentum=0.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, S)
    shadow = [p.detach().clone() for p in p_ref]
    ours = N.FusedRMSpropEMA(p_our, lr=1e-3, alpha=0.999, eps=1e-10, ema_decay=decay, num_iters=S)
    for it in range(S):
        grads = [torch.randn(s, generator=g).cuda() * (10.0 ** ((it % 5) - 3)) for s in shapes]
        for p, q, gr in zip(p_ref, p_our, grads):
            p.grad, q.grad = gr.clone(), gr.clone()
        opt.step()
        sched.step()
        d = min(decay, (1 + it + 1) / (10 + it + 1))
        with torch.no_grad():
            for sh, p in zip(shadow, p_ref):
                sh.sub_((1 - d) * (sh - p))
        ours.step()
    for p, q, sh, sq in zip(p_ref, p_our, shadow, ours.shadow):
        assert rel(q.detach().cpu().numpy(), p.detach().cpu().numpy()) < 1e-6
        assert rel(sq.cpu().numpy(), sh.cpu().numpy()) < 1e-6


def test_device_sampler_statistics_and_reproducibility():
    x = N.sample_gaussian(1 << 20, 16.0, seed=5)
    y = N.sample_gaussian(1 << 20, 16.0, seed=5)
    z = N.sample_gaussian(1 << 20, 16.0, seed=6)
    assert torch.equal(x, y) and not torch.equal(x, z)
    tail = N.sample_gaussian(1 << 10, 16.0, seed=5, offset=(1 << 20) - (1 << 10))
    assert torch.equal(tail, x[-(1 << 10):])                       # counter-based: offset continues the stream
    xs = x.double().cpu().numpy() / 16.0
    assert abs(xs.mean()) < 5e-3 and abs(xs.std() - 1) < 5e-3
    assert abs((xs[:, 0] * xs[:, 1]).mean()) < 5e-3                # coordinates uncorrelated
    assert abs((xs ** 4).mean() - 3.0) < 0.05                      # kurtosis of a normal
    r2 = (xs ** 2).sum(1)
    assert abs(np.mean(r2 < 2 * np.log(2)) - 0.5) < 5e-3           # chi^2_2 median
    # never exactly at the origin (V = -1/r): the radius grid starts at 2.4e-4 sigma
    big = N.sample_gaussian(1 << 26, 1.0, seed=11)
    assert float((big ** 2).sum(1).min()) > 1e-8


def test_device_samplers_of_the_other_importance_modes():
    # main_pde.py:101-118 on the device: Laplace(0, b) per coordinate, uniform on [-s, s)^2
    n = 1 << 20
    lap = N.sample_points(n, "laplacian", 3.0, seed=5)
    assert torch.equal(lap, N.sample_points(n, "laplacian", 3.0, seed=5))
    assert torch.equal(N.sample_points(1 << 10, "laplacian", 3.0, seed=5, offset=n - (1 << 10)), lap[-(1 << 10):])
    v = lap.double().cpu().numpy() / 3.0
    assert abs(v.mean()) < 5e-3 and abs(np.abs(v).mean() - 1) < 5e-3 and abs((v ** 2).mean() - 2) < 2e-2
    assert abs(np.mean(np.abs(v) < np.log(2)) - 0.5) < 5e-3        # median of |x|/b
    assert np.isfinite(v).all() and abs((v[:, 0] * v[:, 1]).mean()) < 1e-2
    uni = N.sample_points(n, "uniform", 2.0, seed=7).double().cpu().numpy() / 2.0
    assert uni.min() > -1 and uni.max() < 1
    assert abs(uni.mean()) < 3e-3 and abs((uni ** 2).mean() - 1 / 3) < 3e-3 and abs((uni[:, 0] * uni[:, 1]).mean()) < 3e-3
    g1, g2 = N.sample_points(4096, "gaussian", 16.0, seed=5), N.sample_gaussian(4096, 16.0, seed=5)
    assert torch.equal(g1, g2)


@pytest.mark.parametrize("engine", ENGINES)
def test_graphed_step_equals_eager_step(engine):
    d, cfg = load_golden("osc_b512_jnt_L16")
    N.set_engine(engine)
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda")
    x = torch.from_numpy(d["x"]).cuda()
    step = N.GraphedOperatorStep(method, operator, importance, batch_size=x.shape[0])
    for rep in range(2):                                        # replay twice: buffers are reused correctly
        loss = step(x)
        grads_g = {n: p.grad.clone() for n, p in method.named_parameters() if p.grad is not None}
        assert abs(float(loss.detach()) - float(d["loss64"])) <= TOL * abs(float(d["loss64"]))
        errs = golden_grad_errors(d, list(grads_g), {k: v.cpu().numpy() for k, v in grads_g.items()})
        assert max(errs.values()) < TOL, errs
    method.zero_grad(set_to_none=True)
    loss_e, aux = method.compute_loss_operator(operator, x, importance=importance)
    loss_e.backward()
    assert abs(float(loss.detach()) - float(loss_e)) < 1e-5 * abs(float(loss_e))
    for n, p in method.named_parameters():
        if p.grad is not None:
            assert rel(grads_g[n].cpu().numpy(), p.grad.cpu().numpy()) < 2e-5, n
    assert rel(step.aux["Tf"].cpu().numpy(), aux["Tf"].cpu().numpy()) < 1e-6


def test_micro_batch_boundaries():
    # several micro-batches with a ragged tail (p_off handling in every kernel): B = 1000 in micro-batches of 256,
    # and B = 65536 + 300 with the default micro-batch, tensor-core engine against the fp32 engine
    d, cfg = load_golden("hyd_small_odd")
    g = torch.Generator().manual_seed(3)
    try:
        for B, mb in ((1000, 256), (65536 + 300, 65536)):
            x = (cfg.sampling_scale * torch.randn(B, 2, generator=g)).cuda()
            res = {}
            for engine in ("fp32", "f16x3"):
                N.set_engine(engine)
                N.set_microbatch(mb)
                method, operator, importance, _ = build_problem(cfg, 21, "cuda")
                loss, aux = method.compute_loss_operator(operator, x, importance=importance)
                loss.backward()
                res[engine] = (float(loss.detach()), aux["Tf"].cpu().numpy(),
                               {n: p.grad.cpu().numpy() for n, p in method.named_parameters() if p.grad is not None})
            assert abs(res["f16x3"][0] - res["fp32"][0]) < TOL * abs(res["fp32"][0])
            assert rel(res["f16x3"][1], res["fp32"][1]) < TOL
            for n in res["fp32"][2]:
                assert rel(res["f16x3"][2][n], res["fp32"][2][n]) < TOL, (B, n)
    finally:
        N.set_microbatch(65536)


def test_config4_shape_L64_engines_agree():
    # BASELINE configs[3] shape (hydrogen, L=64, M_ff=1024) at a batch the fp32 engine finishes quickly:
    # tensor-core engine against the fp32 engine (itself pinned to the L=64 golden fixture above)
    cfg = O.PathConfig.hydrogen(neigs=64)
    g = torch.Generator().manual_seed(64)
    x = (cfg.sampling_scale * torch.randn(3000, 2, generator=g)).cuda()
    res = {}
    for engine in ("fp32", "f16x3"):
        N.set_engine(engine)
        method, operator, importance, _ = build_problem(cfg, 8, "cuda")
        loss, aux = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()
        res[engine] = (float(loss.detach()), aux["f"].cpu().numpy(), aux["Tf"].cpu().numpy(),
                       {n: p.grad.cpu().numpy() for n, p in method.named_parameters() if p.grad is not None})
        del method
        torch.cuda.empty_cache()
    assert abs(res["f16x3"][0] - res["fp32"][0]) < TOL * abs(res["fp32"][0])
    assert rel(res["f16x3"][1], res["fp32"][1]) < TOL and rel(res["f16x3"][2], res["fp32"][2]) < TOL
    for n in res["fp32"][3]:
        assert rel(res["f16x3"][3][n], res["fp32"][3][n]) < TOL, n


@pytest.mark.parametrize("neigs,B", [(64, 65536), (16, 131072)])
def test_full_size_gradient_parity(neigs, B):
    # BASELINE configs[3] at a full micro-batch: hydrogen, L=64, B=65536 (one tensor-core micro-batch: 1024 half tiles
    # per copy in the hidden-layer weight gradient, 64 k-slices in the layer-0 one), and the bench workload itself:
    # L=16, B=131072 (two micro-batches).  Tensor-core engine against the CUDA-core fp32 engine (pinned to the
    # reference at B <= 512 above) on loss, f, Tf and EVERY gradient tensor, and against the numpy oracle on a sample of
    # rows of f / Tf.
    cfg = O.PathConfig.hydrogen(neigs=neigs)
    g = torch.Generator().manual_seed(65)
    x = (cfg.sampling_scale * torch.randn(B, 2, generator=g)).cuda()
    res = {}
    for engine in ("fp32", "f16x3"):
        N.set_engine(engine)
        method, operator, importance, _ = build_problem(cfg, 9, "cuda")
        loss, aux = method.compute_loss_operator(operator, x, importance=importance)
        loss.backward()
        res[engine] = (float(loss.detach()), aux["f"].cpu().numpy(), aux["Tf"].cpu().numpy(),
                       {n: p.grad.cpu().numpy() for n, p in method.named_parameters() if p.grad is not None})
        params = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in method.named_parameters()}
        method.__dict__.pop("_nsvd_scratch", None)
        del method, loss, aux
        torch.cuda.empty_cache()
    tc, ref = res["f16x3"], res["fp32"]
    assert abs(tc[0] - ref[0]) < TOL * abs(ref[0])
    assert rel(tc[1], ref[1]) < TOL and rel(tc[2], ref[2]) < TOL
    errs = {n: rel(tc[3][n], ref[3][n]) for n in ref[3]}
    print(f"L={neigs} B={B} grads tc-vs-fp32:", " ".join(f"{n.split('.')[-2]}{n.split('.')[-1]}={v:.1e}" for n, v in errs.items()))
    assert max(errs.values()) < TOL, errs
    rows = np.random.RandomState(0).choice(B, 1024, replace=False)
    xs = x[torch.from_numpy(rows).cuda()].cpu().numpy().astype(np.float64)
    u = O.forward_streams(xs, params, cfg)
    Tf, f, _ = O.operator_apply(xs, u, params, cfg)
    for eng in (tc, ref):
        assert rel(eng[1][rows], f) < TOL and rel(eng[2][rows], Tf) < TOL


def test_no_grad_and_double_backward_calls():
    d, cfg = load_golden("hyd_small_odd")
    N.set_engine("f16x3")
    method, operator, importance, _ = build_problem(cfg, int(d["seed"]), "cuda")
    x = torch.from_numpy(d["x"]).cuda()
    with torch.no_grad():
        loss0, aux0 = method.compute_loss_operator(operator, x, importance=importance)
    assert not loss0.requires_grad and abs(float(loss0) - float(d["loss64"])) < TOL * abs(float(d["loss64"]))
    loss, _ = method.compute_loss_operator(operator, x, importance=importance)
    loss.backward(retain_graph=True)
    g1 = method.model.base.ws[2].grad.clone()
    loss.backward()                                   # second backward through the same graph accumulates
    assert rel(method.model.base.ws[2].grad.cpu().numpy(), 2 * g1.cpu().numpy()) < 1e-5
    # a second forward invalidates the scratch of the first: its backward must refuse rather than use stale data
    l1, _ = method.compute_loss_operator(operator, x, importance=importance)
    l2, _ = method.compute_loss_operator(operator, x, importance=importance)
    with pytest.raises(RuntimeError, match="overwritten"):
        l1.backward()
    l2.backward()
