"""Randomised agreement of the two arithmetic engines over problem shapes and families: the tcgen05 engine (tiles,
CTA pairs, micro-batches, zero-filled edges) against the fp32 CUDA-core engine (itself pinned to the oracle and the
golden vectors) on shapes no fixture covers."""
import numpy as np
import pytest
import torch

import neural_svd_b200 as N
from conftest import build_problem, rel
from oracle import nsvd_oracle as O

pytestmark = pytest.mark.gpu


def _random_cfg(rs):
    fam = rs.choice(["hydrogen", "oscillator", "well", "cosine", "molion"])
    L = int(rs.choice([2, 3, 7, 16, 17, 31, 40]))
    Mff = int(rs.choice([8, 40, 64, 136, 264]))
    common = dict(neigs=L, fourier_mapping_size=Mff, sequential=bool(rs.randint(2)), step=int(rs.choice([1, 2])))
    if fam == "hydrogen":
        return O.PathConfig.hydrogen(**common)
    if fam == "oscillator":
        return O.PathConfig.oscillator(**common)
    if fam == "well":
        return O.PathConfig(potential="infinite_well", fourier_scale=0.5, operator_scale=1.0, operator_shift=30.0,
                            sampling_mode="uniform", sampling_scale=1.5, lim=1.5, apply_boundary=True,
                            boundary_mode=str(rs.choice(["dir_box_sqrt", "dir_box_exp"])), **common)
    if fam == "cosine":
        return O.PathConfig(potential="cosine", fourier_scale=1.0, fourier_deterministic=True, operator_scale=1.0,
                            operator_shift=2.0, sampling_mode="uniform", sampling_scale=np.pi, lim=np.pi,
                            **{**common, "fourier_mapping_size": Mff // 2 if Mff >= 16 else 8, "neigs": min(L, 25)})
    return O.PathConfig(potential="hydrogen_mol_ion", fourier_scale=0.2, operator_scale=10.0,
                        sampling_mode="laplacian", sampling_scale=3.0, apply_exp_mask=True, exp_mask_init_scale=8.0,
                        hydrogen_mol_ion_R=1.5, **common)


@pytest.mark.parametrize("seed", range(24))
def test_engines_agree_on_random_shapes(seed):
    rs = np.random.RandomState(100 + seed)
    cfg = _random_cfg(rs)
    B = int(rs.choice([2, 3, 127, 129, 255, 1000, 2049, 4097]))
    N.set_microbatch(int(rs.choice([128, 384, 1024, 65536])))
    try:
        g = torch.Generator().manual_seed(seed)
        if cfg.sampling_mode == "uniform":
            x = cfg.sampling_scale * (2 * torch.rand(B, 2, generator=g) - 1)
        else:
            x = cfg.sampling_scale * torch.randn(B, 2, generator=g)
        out = {}
        for engine in ("fp32", "f16x3"):
            N.set_engine(engine)
            method, operator, importance, _ = build_problem(cfg, 50 + seed, "cuda")
            with torch.no_grad():                              # exercise the biases (zero at initialisation)
                gb = torch.Generator().manual_seed(seed)
                for b in method.model.base.bs:
                    b.add_(0.1 * torch.randn(b.shape, generator=gb).to(b.device))
            loss, aux = method.compute_loss_operator(operator, x.cuda(), importance=importance)
            loss.backward()
            out[engine] = (float(loss.detach()), aux["f"].cpu().numpy(), aux["Tf"].cpu().numpy(),
                           {n: p.grad.cpu().numpy() for n, p in method.named_parameters() if p.grad is not None})
        a, b = out["fp32"], out["f16x3"]
        assert np.isfinite(a[0]) and abs(a[0] - b[0]) <= 1e-4 * abs(a[0]), (cfg, B)
        assert rel(b[1], a[1]) < 1e-4 and rel(b[2], a[2]) < 1e-4, (cfg, B)
        assert sorted(a[3]) == sorted(b[3])
        for n in a[3]:
            assert rel(b[3][n], a[3][n]) < 1e-4, (n, cfg, B)
    finally:
        N.set_microbatch(65536)
