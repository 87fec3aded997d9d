"""the fused step accepts the REFERENCE's own objects (its NestedLoRA / WaveFunctions /
OperatorWrapper(NegativeHamiltonian) and the `importance_train` closure over a MultivariateNormal) by duck typing.
Uses the unmodified copy of the reference staged in baseline/_ref (git-ignored, travels to the GPU box); skipped
when that copy is absent."""
import numpy as np
import pytest
import torch

import neural_svd_b200 as N
from conftest import load_golden, rel
from oracle import nsvd_oracle as O
from oracle import ref_bootstrap as RB

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", ["hydrogen", "oscillator", "well_uniform_boxsqrt", "molion_laplace_boxexp_mask",
                                   "cosine_uniform_detff"])
def test_reference_objects_run_on_the_fused_kernels(which):
    root = RB.find_reference()
    if root is None:
        pytest.skip("no copy of the reference (baseline/_ref) on this machine")
    ref = RB.import_reference(root)
    if which == "hydrogen":
        cfg = O.PathConfig.hydrogen(neigs=4, fourier_mapping_size=64, sequential=True)
    elif which == "oscillator":
        cfg = O.PathConfig.oscillator(neigs=4, fourier_mapping_size=64)
    else:                                  # the f-4 families: configuration of the fixture of that name
        cfg = load_golden(which)[1]
    g = torch.Generator().manual_seed(0)
    if cfg.sampling_mode == "uniform":
        x = (cfg.sampling_scale * (2 * torch.rand(200, 1, 2, generator=g) - 1)).reshape(200, -1)
    else:
        x = (cfg.sampling_scale * torch.randn(200, 1, 2, generator=g)).reshape(200, -1)
    # the reference, on the CPU, exact Laplacian
    m_cpu, op_cpu, imp_cpu, _ = RB.build_reference_problem(ref, cfg, 31, 0.0)
    loss_ref, aux_ref = m_cpu.compute_loss_operator(op_cpu, x, importance=imp_cpu)
    loss_ref.backward()
    # the SAME reference classes, parameters moved to the GPU, stepped by our kernels
    m_gpu, op_gpu, imp_gpu, _ = RB.build_reference_problem(ref, cfg, 31, 0.0)
    m_gpu = m_gpu.to("cuda")
    N.set_engine("f16x3")
    loss, aux = N.compute_loss_operator(m_gpu, op_gpu, x.cuda(), imp_gpu)
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) < 1e-4 * abs(float(loss_ref.detach()))
    assert rel(aux["Tf"].cpu().numpy(), aux_ref["Tf"].detach().numpy()) < 1e-4
    got = dict(m_gpu.named_parameters())
    for n, p in m_cpu.named_parameters():
        if p.grad is None:
            assert got[n].grad is None
        else:
            assert rel(got[n].grad.cpu().numpy(), p.grad.numpy()) < 1e-4, n
    assert sorted(m_gpu.state_dict()) == sorted(m_cpu.state_dict())
