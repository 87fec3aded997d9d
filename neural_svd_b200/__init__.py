"""neural_svd_b200 — B200-native (sm_100a) NestedLoRA training step, drop-in for the hot path of
jongharyu/neural-svd (methods/nestedlora.py on the operators of examples/operator + the CDK loss).

Public names mirror the reference's; see DESIGN.md and INTEGRATION.md.
"""
from .models import (DirichletBoundaryMaskBox, ExponentialMask, GaussianFourierFeatureTransform, ParallelMLP, WaveFunctions,
                     get_mlp_eigfuncs, get_wavefunctions)
from .nestedlora import (NestedLoRA, NestedLoRAForCDK, NestedLoRALossFunctionEVD, NestedLoRALossFunctionForCDK,
                         get_joint_nesting_masks, get_sequential_nesting_masks)
from .operators import (GaussianImportance, LaplaceImportance, NegativeHamiltonian, OperatorWrapper,
                        UniformImportance, cosine_potential, get_problem, harmonic_oscillator_potential,
                        hydrogen_mol_ion_potential, hydrogen_potential, infinite_well_potential,
                        make_gaussian_sampler)
from .fused import compute_loss_operator, get_engine, set_engine, set_microbatch
from .dist import PointParallel, shard_points
from .spectrum import compute_spectrum_evd
from .optim import FusedRMSpropEMA, sample_gaussian, sample_points
from .graphs import GraphedOperatorStep
from .siam import HeteroNetwork, get_mlp, get_sketchy_encoder, normalize

__all__ = [n for n in dir() if not n.startswith("_")]
