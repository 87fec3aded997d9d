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
def test_encoder_mirror_equals_reference_modules(mode):
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
