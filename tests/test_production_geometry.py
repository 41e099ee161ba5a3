"""The six attention module classes at PAVE-Net's PRODUCTION geometry (embed_dims 256,
8 heads x 32 channels, 4 levels, P = 4 / 15 / 17;
configs/videopose/2025-2-13/2025_2_13_res50_num_frames_3_posetrack17.py:53-107), forward
and backward, against fixtures produced by executing the reference classes
(tests/golden/gen_golden.py::gen_module_golden_256, tests/golden/module_golden_256.npz).

CPU half (`-m "not gpu"`): the oracle's module compositions (oracle/msda_oracle.py) against
the fixtures — outputs and, through autograd, every gradient.
GPU half (`-m gpu`): this repository's module classes on their default (product) path —
fused-prologue sampling kernels + tcgen05 3xTF32 projections — against the same fixtures,
asserting through the library's per-family launch counters that those kernels are the ones
that ran.

Tolerances: outputs <= 1e-4, gradients <= 1e-3 relative (max|a-b| / max|b|), BASELINE.json.
"""
import os
import sys

import pytest
import torch

from conftest import GOLDEN_DIR, Golden, rel_err
from oracle import msda_oracle as O

sys.path.insert(0, GOLDEN_DIR)
import recipe256  # noqa: E402

CASE_NAMES = [c['name'] for c in recipe256.CASES]

ORACLE_FN = {'encoder': O.encoder_attention_ref, 'pose': O.pose_attention_ref,
             'mf_pose': O.mulframes_pose_attention_ref, 'mf_joint': O.mulframes_joint_attention_ref}


@pytest.fixture(scope='module')
def golden256():
    return Golden(os.path.join(GOLDEN_DIR, 'module_golden_256.npz'))


def _build_cpu_module(case):
    """This repository's module class on the CPU: only used to obtain parameter names /
    shapes for `recipe256.randomise` (construction runs no kernel)."""
    import pavenet_b200
    kw = dict(case['cfg'])
    return getattr(pavenet_b200, case['cls'])(dropout=0.0, **kw).eval()


def _check_grads(tag, case, g, named_grads, leaf_grads, tol):
    for k, v in leaf_grads.items():
        assert rel_err(v, g['grad_in.' + k]) < tol, (tag, 'grad_in', k)
    for pname, grad in named_grads:
        if grad.dim() == 1:
            assert rel_err(grad, g['grad_param.' + pname]) < tol, (tag, pname)
        else:
            u, v = recipe256.projection_vectors(case, pname, grad.shape)
            gc = grad.detach().cpu()
            assert rel_err(u @ gc, g['grad_param_u.' + pname]) < tol, (tag, 'u^T dW', pname)
            assert rel_err(gc @ v, g['grad_param_v.' + pname]) < tol, (tag, 'dW v', pname)


@pytest.mark.parametrize('name', CASE_NAMES)
def test_oracle_compositions_at_production_geometry(golden256, name):
    case = recipe256.case(name)
    g = golden256.case(name)
    mod = _build_cpu_module(case)
    recipe256.randomise(mod, case['param_seed'])
    inp = recipe256.make_inputs(case)
    assert torch.allclose(recipe256.checksum(mod, inp), g['checksum'].double(), rtol=1e-5), \
        'regenerated parameters / inputs differ from the generator run'
    state = {k: v.detach().clone().requires_grad_(True) for k, v in mod.state_dict().items()}
    cfg = dict(case['cfg'], num_frames=case.get('T', 1))
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ('query', 'value', 'query_pos')
              if k in inp}
    kwargs = dict(query_pos=leaves.get('query_pos'), key_padding_mask=inp.get('key_padding_mask'),
                  reference_points=inp['reference_points'], spatial_shapes=inp['spatial_shapes'])
    if case['kind'] == 'encoder':
        out = ORACLE_FN['encoder'](state, cfg, leaves['query'], **kwargs)
    else:
        out = ORACLE_FN[case['kind']](state, cfg, leaves['query'], leaves['value'], **kwargs)
    assert rel_err(out, g['out']) < 1e-5
    out.backward(recipe256.grad_output(case, out.shape))
    _check_grads(name, case, g, [(k, v.grad) for k, v in sorted(state.items())],
                 {k: v.grad for k, v in leaves.items()}, 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('value_dtype', [None, torch.bfloat16])
@pytest.mark.parametrize('name', CASE_NAMES)
def test_product_modules_at_production_geometry(golden256, name, value_dtype):
    import pavenet_b200
    from pavenet_b200 import _capi
    case = recipe256.case(name)
    g = golden256.case(name)
    kw = dict(case['cfg'])
    mod = getattr(pavenet_b200, case['cls'])(dropout=0.0, value_dtype=value_dtype, **kw).eval()
    assert mod.fuse_prologue and mod.tensor_core_linear      # the defaults = the product path
    recipe256.randomise(mod, case['param_seed'])
    inp = recipe256.make_inputs(case)
    assert torch.allclose(recipe256.checksum(mod, inp), g['checksum'].double(), rtol=1e-5)
    mod = mod.cuda()
    dev = {k: v.cuda() for k, v in inp.items()}
    leaves = {k: dev[k].clone().requires_grad_(True) for k in ('query', 'value', 'query_pos')
              if k in dev}
    lsi = O.level_start_index(inp['spatial_shapes']).cuda()
    before = _capi.family_counts()
    out = mod(leaves['query'], None, leaves.get('value'), query_pos=leaves.get('query_pos'),
              key_padding_mask=dev.get('key_padding_mask'),
              reference_points=dev['reference_points'], spatial_shapes=dev['spatial_shapes'],
              level_start_index=lsi)
    out.backward(recipe256.grad_output(case, out.shape).cuda())
    torch.cuda.synchronize()
    after = _capi.family_counts()
    ran = {k: after[k] - before[k] for k in after if after[k] != before[k]}
    # the sampling kernels with the fused prologue (rows or flat family, by problem size) ...
    assert ran.get('fwd_rows_fused', 0) + ran.get('fwd_flat_fused', 0) >= 1, ran
    assert ran.get('bwd_rows_fused', 0) + ran.get('bwd_flat_fused', 0) >= 1, ran
    # ... the tcgen05 projections and their weight gradients, and nothing generic
    assert ran.get('linear', 0) >= 2 and ran.get('linear_wgrad', 0) >= 2, ran
    assert 'fwd_generic' not in ran and 'bwd_generic' not in ran, ran
    if value_dtype is None:
        ftol, btol = 1e-4, 1e-3
    else:
        # bf16 value storage: value rounded to 8 bits of mantissa once (2^-9 relative per
        # element), grad_value rounded once on the way back; stated bound 4e-3 / 8e-3
        ftol, btol = 4e-3, 8e-3
    assert rel_err(out, g['out']) < ftol
    _check_grads(name, case, g, [(k, p.grad) for k, p in sorted(mod.named_parameters())],
                 {k: v.grad for k, v in leaves.items()}, btol)
