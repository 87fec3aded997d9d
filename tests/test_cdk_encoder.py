"""the CDK encoder mirror (neural_svd_b200/siam.py) against the reference's own HeteroNetwork /
get_mlp / normalize (examples/models/siam.py:132-186, mlp.py:129-164), and the whole CDK step (towers -> fused
nsvd_cdk_* loss kernels -> tower gradients) against the reference on the CPU."""
import numpy as np
import pytest
import torch

import neural_svd_b200 as N
from conftest import rel
from oracle import ref_bootstrap as RB


def _reference_encoder(dims, mu, mode, act="lrelu0.2", use_bn=False):
    root = RB.find_reference()
    if root is None:
        pytest.skip("no copy of the reference (baseline/_ref) on this machine")
    RB.import_reference(root)
    from examples.models.mlp import get_mlp
    from examples.models.siam import HeteroNetwork, normalize
    sizes = [dims[0]] + list(dims[1:])
    net = HeteroNetwork(backbones=[get_mlp(sizes=sizes, bias=True, nonlinearity=act, use_bn=use_bn) for _ in range(2)],
                        projectors=[torch.nn.Identity(), torch.nn.Identity()], mu=mu, regularize_mode=mode)
    return net, normalize


@pytest.mark.parametrize("mode", ["l2_ball", "l2_sphere", "clip", "tanh"])
def test_encoder_mirror_equals_reference_modules(mode, monkeypatch):
    # module structure / state-dict / normalize on the CPU: the dense layers run their host mirror here (the product
    # path refuses CPU tensors, see test_dense_layer_has_no_cpu_path)
    from neural_svd_b200 import linear
    monkeypatch.setattr(linear, "HOST_MIRROR", True)
    torch.manual_seed(0)
    ref, ref_normalize = _reference_encoder([48, 96, 32], 4.0, mode)
    mine = N.get_sketchy_encoder(network_dims="96,32", mu=4.0, regularize_mode=mode, input_dim=48)
    assert list(mine.state_dict()) == list(ref.state_dict())          # checkpoints are interchangeable
    mine.load_state_dict(ref.state_dict())
    x, y = 3 * torch.randn(40, 48), 3 * torch.randn(40, 48)
    out_r, out_m = ref(x, y), mine(x, y)
    for a, b in zip(out_r, out_m):
        assert torch.equal(a, b)
    (out_r[1].square().sum() + out_r[3].sum()).backward()
    (out_m[1].square().sum() + out_m[3].sum()).backward()
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        assert torch.equal(p.grad, q.grad), n
    z = 2 * torch.randn(16, 8)
    assert torch.equal(N.normalize(z, 2.0, mode), ref_normalize(z, 2.0, mode))
    assert mine.output_dims == ref.output_dims


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["fp32", "f16x3"])
def test_cdk_step_with_encoder_matches_reference(engine):
    torch.manual_seed(1)
    ref_net, _ = _reference_encoder([64, 256, 48], 16.0, "l2_ball")
    ref = RB.import_reference(RB.find_reference())
    B, L = 512, 48
    x, y = torch.randn(B, 64), torch.randn(B, 64)
    # the reference, CPU fp32: towers + NestedLoRAForCDK (methods/nestedlora.py:335-378)
    m_ref = ref.NestedLoRAForCDK(model=ref_net, neigs=L, step=1, sequential=False, set_first_mode_const=True)
    _, fx, _, fy = m_ref(x, y)
    loss_ref, *_ = m_ref.compute_loss(fx, fy)
    loss_ref.backward()
    # this package on the GPU: same weights, fused CDK loss kernels
    net = N.get_sketchy_encoder(network_dims="256,48", mu=16.0, input_dim=64)
    net.load_state_dict(ref_net.state_dict())
    N.set_engine(engine)
    m = N.NestedLoRAForCDK(model=net, neigs=L, step=1, sequential=False, set_first_mode_const=True).cuda()
    _, gx, _, gy = m(x.cuda(), y.cuda())
    loss, lop, lmet, rs_joint, rs_indep = m.compute_loss(gx, gy)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < 1e-4 * abs(float(loss_ref.detach()))
    assert rs_joint.shape == (B,) and rs_indep.shape == (B * B - B,)
    for (n, p), (_, q) in zip(ref_net.named_parameters(), net.named_parameters()):
        assert rel(q.grad.cpu().numpy(), p.grad.numpy()) < 1e-4, n


def test_dense_layer_has_no_cpu_path():
    from neural_svd_b200.linear import TCLinear
    with pytest.raises(RuntimeError, match="no CPU path"):
        TCLinear(16, 8)(torch.randn(4, 16))


def test_get_mlp_structure_matches_reference_indices():
    """fused activations keep their slot in the Sequential, so `backbones.x.<i>.weight` indices do not move"""
    from neural_svd_b200.linear import FusedActivation, TCLinear
    m = N.get_mlp([8, 16, 24, 8], nonlinearity="lrelu0.2", use_bn=False)
    assert [type(l) for l in m] == [TCLinear, FusedActivation, TCLinear, FusedActivation, TCLinear]
    assert m[0].fused_act == ("leaky", 0.2) and m[4].fused_act is None
    m = N.get_mlp([8, 16, 8], nonlinearity="relu", use_bn=True)      # BatchNorm sits between Linear and ReLU: not fused
    assert [type(l).__name__ for l in m] == ["TCLinear", "BatchNorm1d", "ReLU", "TCLinear", "BatchNorm1d"]
    m = N.get_mlp([8, 16, 8], nonlinearity="tanh", use_bn=False)
    assert [type(l).__name__ for l in m] == ["TCLinear", "Tanh", "TCLinear"]


@pytest.mark.gpu
@pytest.mark.parametrize("rows,in_f,out_f,act", [(300, 72, 200, ("leaky", 0.2)), (256, 64, 48, None), (1, 8, 8, ("leaky", 0.0)),
                                                 (1000, 520, 136, ("leaky", 0.2))])
def test_dense_layer_kernels_match_fp64(rows, in_f, out_f, act):
    """nsvd_linear_fwd / nsvd_linear_bwd (ragged shapes: partial tiles in every dimension) against fp64 torch"""
    from neural_svd_b200.linear import TCLinear
    torch.manual_seed(rows + in_f)
    lin = TCLinear(in_f, out_f, fused_act=act).cuda()
    x = (2 * torch.randn(rows, in_f)).cuda().requires_grad_()
    gy = torch.randn(rows, out_f).cuda()
    y = lin(x)
    y.backward(gy)
    xd = x.detach().double().requires_grad_()
    wd, bd = lin.weight.detach().double().requires_grad_(), lin.bias.detach().double().requires_grad_()
    yd = torch.nn.functional.linear(xd, wd, bd)
    if act is not None:
        yd = torch.where(y.detach() > 0, yd, act[1] * yd)
    yd.backward(gy.double())
    # the activation mask of the fp64 reference is taken from OUR y: the kink is not what this test is about
    assert rel(y.detach().cpu().numpy(), yd.detach().cpu().numpy()) < 2e-6
    assert rel(x.grad.cpu().numpy(), xd.grad.cpu().numpy()) < 2e-6
    assert rel(lin.weight.grad.cpu().numpy(), wd.grad.cpu().numpy()) < 2e-6
    assert rel(lin.bias.grad.cpu().numpy(), bd.grad.cpu().numpy()) < 2e-6


@pytest.mark.gpu
def test_sketchy_towers_full_size_against_cublas_fp32():
    """config-5 encoder shape (4096 x 512 -> 8192 -> 512, lrelu0.2, l2_ball): hand-written kernels vs torch/cuBLAS fp32
    (TF32 off) with the same weights; sampled fp64 check of the first layer.  Behind the LeakyReLU kink a hidden
    activation within rounding distance of zero may take the other slope in either implementation (33 M activations per
    tower: a handful do, also between cuBLAS and the CPU); the rows of the first layer's weight gradient that belong to
    such hidden units are compared separately, everything else at 5e-5."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    net = N.get_sketchy_encoder().cuda()
    x, y = torch.randn(4096, 512, device="cuda"), torch.randn(4096, 512, device="cuda")
    _, ex, _, ey = net(x, y)
    (ex.square().sum() + (ex * ey).sum()).backward()
    ref = torch.nn.Sequential(torch.nn.Linear(512, 8192), torch.nn.LeakyReLU(0.2), torch.nn.Linear(8192, 512)).cuda()
    ref.load_state_dict(net.backbones["x"].state_dict())
    rx = N.normalize(ref(x), 4.0, "l2_ball")
    assert rel(ex.detach().cpu().numpy(), rx.detach().cpu().numpy()) < 2e-5
    (rx.square().sum() + (rx * ey.detach()).sum()).backward()
    with torch.no_grad():
        h_mine, h_ref = net.backbones["x"][0](x), ref[1](ref[0](x))
        flipped = (h_mine > 0) != (h_ref > 0)
    n_flips = int(flipped.sum())
    clean = (~flipped.any(dim=0)).cpu().numpy()          # hidden units without a flipped activation
    print("kink flips:", n_flips, "of", flipped.numel(), "| hidden units touched:", int((~clean).sum()))
    assert n_flips <= 32
    g = {n: (p.grad.cpu().numpy(), q.grad.cpu().numpy())
         for (n, p), (_, q) in zip(net.backbones["x"].named_parameters(), ref.named_parameters())}
    assert rel(g["0.weight"][0][clean], g["0.weight"][1][clean]) < 5e-5
    assert rel(g["0.bias"][0][clean], g["0.bias"][1][clean]) < 5e-5
    assert rel(g["0.weight"][0], g["0.weight"][1]) < 2e-3     # with the flipped units: one row each, O(1/64) relative
    assert rel(g["2.weight"][0], g["2.weight"][1]) < 5e-5
    assert rel(g["2.bias"][0], g["2.bias"][1]) < 5e-5
    idx = torch.randint(0, 4096, (64,), device="cuda")
    h64 = torch.nn.functional.leaky_relu(x[idx].double() @ net.backbones["x"][0].weight.double().T
                                         + net.backbones["x"][0].bias.double(), 0.2)
    assert rel(h_mine[idx].cpu().numpy(), h64.detach().cpu().numpy()) < 2e-6
