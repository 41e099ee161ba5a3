"""Pins the CPU oracle (oracle/) to the golden vectors produced by executing
the reference (tests/golden/gen_golden.py).  CPU only."""
import os

import pytest
import torch

from conftest import rel_err
from oracle import msda_oracle as O

OP_CASES = ['mmcv_f64', 'mmcv_f32', 'gradcheck_c4', 'gradcheck_c30', 'gradcheck_c32',
            'gradcheck_c64', 'gradcheck_c71', 'gradcheck_c1025', 'enc_f32', 'enc_f64',
            'pose17_f32', 'pose15_f64', 'd16_f32', 'd64_f32', 'edge_f32', 'edge_f64']


def _tol(dtype):
    return 2e-6 if dtype == torch.float32 else 1e-13


def _grad_queries(name, t):
    """Exact-border / exact-pixel-centre locations are points where the
    bilinear surface has a kink: the location gradient is one-sided there, so
    for the hand-placed edge cases only the nudged queries are compared."""
    return slice(1, None) if name.startswith('edge') else slice(None)


@pytest.mark.parametrize('name', OP_CASES)
def test_c_oracle_matches_reference(op_golden, name):
    c = op_golden.case(name)
    lsi = O.level_start_index(c['shapes'])
    tol = _tol(c['loc'].dtype)
    out = O.c_forward(c['value'], c['shapes'], lsi, c['loc'], c['aw'])
    assert out.dtype == c['out'].dtype
    assert rel_err(out, c['out']) < tol
    gv, gl, ga = O.c_backward(c['value'], c['shapes'], lsi, c['loc'], c['aw'], c['grad_out'])
    q = _grad_queries(name, c)
    assert rel_err(gv[:, :], c['grad_value']) < tol or name.startswith('edge')
    assert rel_err(gl[:, q], c['grad_loc'][:, q]) < tol
    assert rel_err(ga, c['grad_aw']) < tol


@pytest.mark.parametrize('name', OP_CASES)
def test_grid_sample_port_matches_reference(op_golden, name):
    c = op_golden.case(name)
    v = c['value'].clone().requires_grad_()
    loc = c['loc'].clone().requires_grad_()
    aw = c['aw'].clone().requires_grad_()
    out = O.grid_sample_port(v, c['shapes'], loc, aw)
    out.backward(c['grad_out'])
    tol = _tol(loc.dtype)
    assert rel_err(out, c['out']) < tol
    assert rel_err(v.grad, c['grad_value']) < tol
    assert rel_err(loc.grad, c['grad_loc']) < tol
    assert rel_err(aw.grad, c['grad_aw']) < tol


def test_mmcv_float_tolerances(op_golden):
    """The bounds of test_forward_equal_with_pytorch_float
    (test_ms_deformable_attn.py:106-135), applied to the C oracle in fp32."""
    c = op_golden.case('mmcv_f32')
    out = O.c_forward(c['value'], c['shapes'], None, c['loc'], c['aw'])
    assert torch.allclose(out, c['out'], rtol=1e-2, atol=1e-3)
    assert (out - c['out']).abs().max() < 1e-9
    assert ((out - c['out']).abs() / c['out'].abs()).max() < 1e-6


def test_nonfinite_locations_are_skipped(op_golden):
    """NaN / Inf locations fail the range test and contribute nothing
    (SURVEY.md Appendix A).  The clean output is the reference's; poisoning
    locations whose weight we zero must leave it unchanged."""
    c = op_golden.case('nonfinite_f32')
    loc, aw = c['loc'].clone(), c['aw'].clone()
    loc[0, 0, 0, 0, 0, 0] = float('nan')
    loc[0, 1, 1, 1, 2, 1] = float('inf')
    loc[0, 2, 0, 3, 1, 0] = -float('inf')
    out = O.c_forward(c['value'], c['shapes'], None, loc, aw)
    assert torch.isfinite(out).all()
    # expected = clean output minus the contributions of the poisoned samples
    aw0 = aw.clone()
    aw0[0, 0, 0, 0, 0] = 0
    aw0[0, 1, 1, 1, 2] = 0
    aw0[0, 2, 0, 3, 1] = 0
    expect = O.grid_sample_port(c['value'], c['shapes'], c['loc'], aw0)
    assert rel_err(out, expect) < 2e-6


def _module_args(c):
    state = {k[len('state.'):]: v for k, v in c.items() if k.startswith('state.')}
    cfg = {k[len('cfg.'):]: int(v) for k, v in c.items() if k.startswith('cfg.')}
    inp = {k[len('in.'):]: v for k, v in c.items() if k.startswith('in.')}
    return state, cfg, inp


def test_encoder_composition(module_golden):
    for name in ('encoder', 'encoder_box'):
        state, cfg, inp = _module_args(module_golden.case(name))
        out = O.encoder_attention_ref(
            state, cfg, inp['query'], value=inp.get('value'), query_pos=inp.get('query_pos'),
            key_padding_mask=inp.get('key_padding_mask'),
            reference_points=inp['reference_points'], spatial_shapes=inp['spatial_shapes'])
        assert rel_err(out, module_golden.case(name)['out']) < 1e-5


def test_pose_composition(module_golden):
    state, cfg, inp = _module_args(module_golden.case('pose'))
    out = O.pose_attention_ref(state, cfg, inp['query'], inp['value'], query_pos=inp['query_pos'],
                               key_padding_mask=inp['key_padding_mask'],
                               reference_points=inp['reference_points'],
                               spatial_shapes=inp['spatial_shapes'])
    assert rel_err(out, module_golden.case('pose')['out']) < 1e-5


@pytest.mark.parametrize('T', [3, 5])
def test_mulframes_pose_composition(module_golden, T):
    c = module_golden.case('mf_pose%d' % T)
    state, cfg, inp = _module_args(c)
    out = O.mulframes_pose_attention_ref(
        state, cfg, inp['query'], inp['value'], query_pos=inp['query_pos'],
        key_padding_mask=inp['key_padding_mask'], reference_points=inp['reference_points'],
        spatial_shapes=inp['spatial_shapes'])
    assert rel_err(out, c['out']) < 1e-5


@pytest.mark.parametrize('T', [3, 5])
def test_mulframes_joint_composition(module_golden, T):
    c = module_golden.case('mf_joint%d' % T)
    state, cfg, inp = _module_args(c)
    out = O.mulframes_joint_attention_ref(
        state, cfg, inp['query'], inp['value'], query_pos=inp['query_pos'],
        key_padding_mask=inp['key_padding_mask'], reference_points=inp['reference_points'],
        spatial_shapes=inp['spatial_shapes'])
    assert rel_err(out, c['out']) < 1e-5


def test_c_oracle_agrees_with_port_on_random_shapes():
    """Two independent restatements (scalar C loops vs grid_sample) on shapes
    not in the fixtures, including ragged level sizes."""
    g = torch.Generator().manual_seed(77)
    for shapes, M, D, Q, P in (([(7, 3), (1, 9), (5, 5)], 3, 5, 13, 3),
                               ([(1, 1)], 1, 1, 1, 1),
                               ([(13, 21), (25, 42)], 8, 32, 10, 15)):
        shapes = torch.tensor(shapes)
        S = int(shapes.prod(1).sum())
        L = shapes.shape[0]
        v = torch.randn(2, S, M, D, generator=g, dtype=torch.float64)
        loc = torch.rand(2, Q, M, L, P, 2, generator=g, dtype=torch.float64) * 1.4 - 0.2
        aw = torch.rand(2, Q, M, L, P, generator=g, dtype=torch.float64)
        go = torch.randn(2, Q, M * D, generator=g, dtype=torch.float64)
        vv, ll, aa = v.clone().requires_grad_(), loc.clone().requires_grad_(), aw.clone().requires_grad_()
        ref = O.grid_sample_port(vv, shapes, ll, aa)
        ref.backward(go)
        assert rel_err(O.c_forward(v, shapes, None, loc, aw), ref) < 1e-13
        gv, gl, ga = O.c_backward(v, shapes, None, loc, aw, go)
        assert rel_err(gv, vv.grad) < 1e-13
        assert rel_err(gl, ll.grad) < 1e-12
        assert rel_err(ga, aa.grad) < 1e-13


def test_c_oracle_structural_properties():
    """Properties the full-size GPU tests lean on, held by the oracle itself: the op is linear in
    `value` and in the attention weights, blind to the order of the points inside a level, and a
    sample whose weight is zero contributes nothing to the output or to grad_value."""
    g = torch.Generator().manual_seed(5)
    shapes = torch.tensor([(9, 7), (4, 5), (2, 3)])
    S, L, M, D, Q, P = int(shapes.prod(1).sum()), 3, 2, 6, 11, 4
    dt = torch.float64
    v1 = torch.randn(1, S, M, D, generator=g, dtype=dt)
    v2 = torch.randn(1, S, M, D, generator=g, dtype=dt)
    loc = torch.rand(1, Q, M, L, P, 2, generator=g, dtype=dt) * 1.3 - 0.15
    aw = torch.rand(1, Q, M, L, P, generator=g, dtype=dt)
    go = torch.randn(1, Q, M * D, generator=g, dtype=dt)
    f = lambda v, a=aw, l=loc: O.c_forward(v, shapes, None, l, a)   # noqa: E731
    assert rel_err(f(2.5 * v1 - v2), 2.5 * f(v1) - f(v2)) < 1e-13
    assert rel_err(f(v1, 3.0 * aw), 3.0 * f(v1)) < 1e-13
    perm = torch.tensor([2, 0, 3, 1])
    assert rel_err(f(v1, aw[..., perm], loc[..., perm, :]), f(v1)) < 1e-13
    aw0 = aw.clone()
    aw0[..., 1, :] = 0                                  # silence level 1
    v_mod = v1.clone()
    start1 = int(shapes[0].prod())
    v_mod[:, start1:start1 + int(shapes[1].prod())] = 123.0     # ... then its values cannot matter
    assert rel_err(f(v_mod, aw0), f(v1, aw0)) < 1e-13
    gv, _, _ = O.c_backward(v1, shapes, None, loc, aw0, go)
    assert float(gv[:, start1:start1 + int(shapes[1].prod())].abs().max()) == 0.0


@pytest.mark.skipif(not os.path.isdir('/root/reference/third_party/mmcv'),
                    reason='needs the read-only reference checkout (build container only)')
def test_golden_fixtures_regenerate_from_the_reference(tmp_path):
    """The committed fixtures ARE the reference's outputs: re-running the generator against the
    reference checkout (where it exists) reproduces all three .npz files bit for bit."""
    import shutil
    import subprocess
    import sys
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    script = shutil.copy(os.path.join(here, 'gen_golden.py'), str(tmp_path))
    shutil.copy(os.path.join(here, 'recipe256.py'), str(tmp_path))
    proc = subprocess.run([sys.executable, script, '/root/reference'], capture_output=True, text=True,
                          timeout=900, cwd=str(tmp_path))
    assert proc.returncode == 0, proc.stderr[-2000:]
    for name in ('op_golden.npz', 'module_golden.npz', 'module_golden_256.npz'):
        new = np.load(os.path.join(str(tmp_path), name), allow_pickle=True)
        old = np.load(os.path.join(here, name), allow_pickle=True)
        assert sorted(new.files) == sorted(old.files)
        for k in new.files:
            assert new[k].dtype == old[k].dtype and new[k].shape == old[k].shape, k
            assert np.array_equal(new[k], old[k], equal_nan=new[k].dtype.kind == 'f'), k
