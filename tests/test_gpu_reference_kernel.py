"""The new kernels against the REFERENCE'S OWN CUDA kernels on the same GPU and tensors.

oracle/_ref/libmsda_refcuda.so is the reference's ms_deform_attn_cuda_kernel.cuh compiled for
sm_100a from the reference checkout (oracle/Makefile `refcuda`, oracle/refcuda_driver.cu) — the
implementation PAVE-Net actually trains with, as opposed to its CPU fallback.  Both sides
compute in fp32 with different summation orders, so they agree to rounding:
outputs <= 1e-5, gradients <= 1e-4 (max|a-b| / max|b|); the contract is 1e-4 / 1e-3.
"""
import pytest
import torch

from conftest import rel_err
from oracle import msda_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not O.refcuda_available(),
                                 reason='oracle/_ref/libmsda_refcuda.so not built (needs the '
                                        'reference checkout: make -C oracle refcuda)')]

MID = [(28, 40), (14, 20), (7, 10), (4, 5)]


def _problem(seed, B, Q, M, D, P, shapes, dtype=torch.float32, spread=0.15):
    g = torch.Generator().manual_seed(seed)
    shapes_t = torch.tensor(shapes, dtype=torch.long)
    L = len(shapes)
    S = int(shapes_t.prod(1).sum())
    value = torch.randn(B, S, M, D, generator=g, dtype=dtype)
    loc = torch.rand(B, Q, M, L, P, 2, generator=g, dtype=dtype) * (1 + 2 * spread) - spread
    aw = torch.softmax(torch.randn(B, Q, M, L * P, generator=g, dtype=dtype), -1).view(B, Q, M, L, P)
    go = torch.randn(B, Q, M * D, generator=g, dtype=dtype)
    return [t.cuda() for t in (value, shapes_t, O.level_start_index(shapes_t), loc, aw, go)]


@pytest.mark.parametrize('B,Q,M,D,P,shapes,dtype', [
    (2, 1500, 8, 32, 4, MID, torch.float32),          # encoder-like (rows family)
    (1, 300, 8, 32, 17, MID * 5, torch.float32),      # 5-frame pose decoder (flat family)
    (1, 300, 8, 32, 17, MID, torch.float32),          # PETR
    (1, 77, 4, 16, 3, MID[:2], torch.float32),
    (1, 41, 2, 64, 5, MID[1:], torch.float32),        # reference picks its _v2<64> backward
    (1, 33, 2, 24, 5, MID, torch.float32),            # not a power of two: shm_reduce_v1 vs generic
    (1, 50, 2, 32, 4, MID, torch.float64),
])
def test_new_kernels_match_the_reference_cuda_kernels(B, Q, M, D, P, shapes, dtype):
    import pavenet_b200
    value, shapes_t, lsi, loc, aw, go = _problem(B * 100 + Q + D, B, Q, M, D, P, shapes, dtype)
    ref_out = O.refcuda_forward(value, shapes_t, lsi, loc, aw)
    rgv, rgl, rga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(aw)
    O.refcuda_backward(value, shapes_t, lsi, loc, aw, go, rgv, rgl, rga)
    v, l, a = (t.clone().requires_grad_() for t in (value, loc, aw))
    out = pavenet_b200.MultiScaleDeformableAttnFunction.apply(v, shapes_t, lsi, l, a, 64)
    out.backward(go)
    torch.cuda.synchronize()
    ftol, btol = (1e-5, 1e-4) if dtype == torch.float32 else (1e-12, 1e-12)
    assert rel_err(out, ref_out) < ftol
    assert rel_err(v.grad, rgv) < btol
    assert rel_err(l.grad, rgl) < btol
    assert rel_err(a.grad, rga) < btol


def test_reference_cuda_kernels_match_the_c_oracle():
    """Closes the triangle: reference CUDA kernels == C restatement (which is pinned to the
    reference's CPU function by tests/test_oracle.py)."""
    value, shapes_t, lsi, loc, aw, go = _problem(9, 2, 200, 8, 32, 4, MID)
    ref_out = O.refcuda_forward(value, shapes_t, lsi, loc, aw)
    rgv, rgl, rga = torch.zeros_like(value), torch.zeros_like(loc), torch.zeros_like(aw)
    O.refcuda_backward(value, shapes_t, lsi, loc, aw, go, rgv, rgl, rga)
    cpu = [t.cpu() for t in (value, shapes_t, lsi, loc, aw)]
    assert rel_err(ref_out, O.c_forward(*cpu)) < 1e-5
    cgv, cgl, cga = O.c_backward(*cpu, go.cpu())
    assert rel_err(rgv, cgv) < 1e-4 and rel_err(rgl, cgl) < 1e-4 and rel_err(rga, cga) < 1e-4
