"""CPU-only tests: the C-ABI library loads and exports what the header
declares, argument validation works without a GPU, and the host-side mirror of
the reference's module interface behaves like the reference."""
import ctypes
import math
import os
import re

import pytest
import torch

import pavenet_b200
from pavenet_b200 import _build, _capi, modules
from oracle import msda_oracle as O
from conftest import ROOT, rel_err


# ---------------------------------------------------------------- C ABI ----
def _declared_functions():
    text = open(os.path.join(ROOT, 'include', 'pavenet_msda.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(msda_[a-z_0-9]+)\s*\(', text)))


def test_library_is_built_in_tree():
    path = _build.build()
    assert os.path.exists(path)
    assert os.path.relpath(path, ROOT).startswith(os.path.join('pavenet_b200', 'lib'))


def test_library_exports_every_declared_symbol():
    lib = _capi.load()
    declared = _declared_functions()
    assert 'msda_forward' in declared and 'msda_backward' in declared
    for name in declared:
        assert hasattr(lib, name), 'missing export: ' + name
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared
    assert lib.msda_abi_version() == 2


def test_library_has_sm100a_code():
    import subprocess
    out = subprocess.run(['cuobjdump', '--list-elf', _build.LIB_PATH],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    assert 'sm_100a' in out, out


def test_argument_validation_needs_no_gpu():
    lib = _capi.load()
    one = ctypes.c_void_p(16)
    # NULL pointer
    rc = lib.msda_forward(None, one, one, one, one, one, 1, 1, 1, 32, 1, 1, 1, 0, 0, None)
    assert rc == -1 and b'NULL' in lib.msda_last_error()
    # non-positive size
    rc = lib.msda_forward(one, one, one, one, one, one, 1, 1, 0, 32, 1, 1, 1, 0, 0, None)
    assert rc == -1 and b'positive' in lib.msda_last_error()
    # bad dtype combination: bf16 value with f64 locations
    rc = lib.msda_forward(one, one, one, one, one, one, 1, 1, 1, 32, 1, 1, 1, 1, 2, None)
    assert rc == -1 and b'value_dtype' in lib.msda_last_error()
    rc = lib.msda_backward(one, one, one, one, one, one, one, one, one,
                           1, 1, 1, 32, 1, 1, 1, 0, 0, 1, None)
    assert rc == -1 and b'grad_value_dtype' in lib.msda_last_error()
    # offsets inside a batch entry must fit int32
    rc = lib.msda_forward(one, one, one, one, one, one, 1, 1 << 24, 8, 32, 1, 1, 1, 0, 0, None)
    assert rc == -2
    with pytest.raises(RuntimeError, match='status -1'):
        _capi.check(-1, 'msda_forward')


def test_kernel_dispatch_names():
    assert _capi.kernel_name(32, _capi.MSDA_F32, _capi.MSDA_F32) == 'rows<D=32,f32>'
    assert _capi.kernel_name(32, _capi.MSDA_F32, _capi.MSDA_BF16) == 'rows<D=32,bf16>'
    assert _capi.kernel_name(32, _capi.MSDA_F64, _capi.MSDA_F64) == 'generic'
    assert _capi.kernel_name(30, _capi.MSDA_F32, _capi.MSDA_F32) == 'generic'


def test_no_cpu_fallback():
    """CPU tensors are refused by the op and by the modules (no silent PyTorch path)."""
    shapes = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    v = torch.randn(1, 4, 1, 32)
    loc = torch.rand(1, 3, 1, 1, 2, 2)
    aw = torch.rand(1, 3, 1, 1, 2)
    with pytest.raises(RuntimeError, match='CUDA tensor'):
        pavenet_b200.MultiScaleDeformableAttnFunction.apply(v, shapes, lsi, loc, aw, 64)
    m = pavenet_b200.MultiScaleDeformableAttention(embed_dims=32, num_heads=1, num_levels=1)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.randn(3, 1, 32), reference_points=torch.rand(1, 3, 1, 2), spatial_shapes=shapes,
          level_start_index=lsi, value=torch.randn(4, 1, 32))


def test_product_code_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pavenet_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f), encoding='utf-8').read()
                hits = re.findall(r'(?:import|from)\s+oracle|msda_oracle|msda_ref|oracle[/\\]',
                                  text)
                assert not hits, '%s reaches into the oracle: %s' % (os.path.join(dirpath, f), hits)


# ------------------------------------------------------- module mirror ----
def test_constructor_errors_and_registry():
    # pinned by the reference's test (test_ms_deformable_attn.py:26-31)
    with pytest.raises(ValueError):
        pavenet_b200.MultiScaleDeformableAttention(embed_dims=256, num_heads=7)
    for t in ('MultiScaleDeformableAttention', 'mmcv.MultiScaleDeformableAttention',
              'opera.MultiScaleDeformablePoseAttention', 'MultiScaleDeformablePoseAttention',
              'opera.MulFramesMultiScaleDeformablePoseAttentionNumFrames5',
              'mmcv.MulFramesMultiScaleDeformableAttentionNumFrames5'):
        assert type(pavenet_b200.build_attention(dict(type=t))).__name__ == t.split('.')[-1]
    # the canonical config entries (configs/videopose/2025-2-13/...posetrack17.py)
    m = pavenet_b200.build_attention(dict(
        type='opera.MulFramesMultiScaleDeformablePoseAttentionNumFrames3', num_frames=3,
        embed_dims=256, num_heads=8, num_levels=4, num_points=15, im2col_step=128))
    assert m.num_points == 15 and m.im2col_step == 128 and m.num_frames == 3
    j = pavenet_b200.build_attention(dict(
        type='mmcv.MulFramesMultiScaleDeformableAttentionNumFrames3', num_frames=3,
        embed_dims=256, num_levels=4, im2col_step=128))
    assert j.num_points == 4
    with pytest.raises(KeyError):
        pavenet_b200.build_attention(dict(type='opera.NoSuchAttention'))


def test_state_dict_keys_match_reference(module_golden):
    """Parameter names are the checkpoint-compatibility contract."""
    for case, cls in (('encoder', 'MultiScaleDeformableAttention'),
                      ('pose', 'MultiScaleDeformablePoseAttention'),
                      ('mf_pose3', 'MulFramesMultiScaleDeformablePoseAttentionNumFrames3'),
                      ('mf_pose5', 'MulFramesMultiScaleDeformablePoseAttentionNumFrames5'),
                      ('mf_joint3', 'MulFramesMultiScaleDeformableAttentionNumFrames3'),
                      ('mf_joint5', 'MulFramesMultiScaleDeformableAttentionNumFrames5')):
        c = module_golden.case(case)
        ref_keys = sorted(k[len('state.'):] for k in c if k.startswith('state.'))
        cfg = {k[len('cfg.'):]: int(v) for k, v in c.items() if k.startswith('cfg.')}
        if cls.endswith('5'):
            cfg.pop('num_frames', None)
        mod = getattr(pavenet_b200, cls)(**cfg)
        assert sorted(mod.state_dict().keys()) == ref_keys, cls
        for k in ref_keys:
            assert tuple(mod.state_dict()[k].shape) == tuple(c['state.' + k].shape)


def test_initialisation_follows_reference():
    m = pavenet_b200.MultiScaleDeformableAttention()
    assert m.sampling_offsets.weight.abs().sum() == 0
    bias = m.sampling_offsets.bias.view(8, 4, 4, 2)
    # head 0 looks along +x, point p at p+1 pixels; head 2 along +y
    assert torch.allclose(bias[0, :, :, 0], torch.arange(1., 5.).expand(4, 4))
    assert bias[0, :, :, 1].abs().max() < 1e-6
    assert torch.allclose(bias[2, 0, :, 1], torch.arange(1., 5.))
    assert m.attention_weights.weight.abs().sum() == 0 and m.attention_weights.bias.abs().sum() == 0
    bound = math.sqrt(6.0 / (256 + 256))
    assert m.value_proj.weight.abs().max() <= bound and m.value_proj.bias.abs().sum() == 0
    p = pavenet_b200.MulFramesMultiScaleDeformablePoseAttentionNumFrames5()
    assert p.num_frames == 5 and p.pre_pre_sampling_offsets.bias.abs().sum() == 0
    j = pavenet_b200.MulFramesMultiScaleDeformableAttentionNumFrames3()
    assert torch.equal(j.pre_sampling_offsets.bias, m.sampling_offsets.bias)
    assert torch.equal(j.next_sampling_offsets.bias, m.sampling_offsets.bias)


def test_fuse_frames_as_levels():
    shapes = torch.tensor([[4, 6], [2, 3]])
    lsi = torch.tensor([0, 24])
    s, i = pavenet_b200.fuse_frames_as_levels(shapes, lsi, 3, 30)
    assert s.tolist() == [[4, 6], [2, 3]] * 3
    assert i.tolist() == [0, 24, 30, 54, 60, 84]


def _with_oracle_op(monkeypatch):
    """Run the module classes on the CPU with the op replaced by the ORACLE —
    test-only, to check the host-side maths around the op without a GPU."""
    def fake(value, shapes, lsi, loc, w, step):
        return O.c_forward(value.float(), shapes, lsi, loc, w).to(loc.dtype)
    monkeypatch.setattr(modules, '_run_op', fake)


def _module_args(c):
    state = {k[len('state.'):]: v for k, v in c.items() if k.startswith('state.')}
    cfg = {k[len('cfg.'):]: int(v) for k, v in c.items() if k.startswith('cfg.')}
    inp = {k[len('in.'):]: v for k, v in c.items() if k.startswith('in.')}
    return state, cfg, inp


@pytest.mark.parametrize('case,cls', [
    ('encoder', 'MultiScaleDeformableAttention'),
    ('encoder_box', 'MultiScaleDeformableAttention'),
    ('pose', 'MultiScaleDeformablePoseAttention'),
    ('mf_pose3', 'MulFramesMultiScaleDeformablePoseAttentionNumFrames3'),
    ('mf_pose5', 'MulFramesMultiScaleDeformablePoseAttentionNumFrames5'),
    ('mf_joint3', 'MulFramesMultiScaleDeformableAttentionNumFrames3'),
    ('mf_joint5', 'MulFramesMultiScaleDeformableAttentionNumFrames5'),
])
@pytest.mark.parametrize('fused', [True, False])
def test_host_side_maths_matches_reference_modules(monkeypatch, module_golden, case, cls, fused):
    """Everything around the op (projections, softmax, reference-point
    transforms, the fused T*L-level reformulation) against the outputs of the
    reference's module classes."""
    _with_oracle_op(monkeypatch)
    c = module_golden.case(case)
    state, cfg, inp = _module_args(c)
    if cls.endswith('5'):
        cfg.pop('num_frames', None)
    extra = {'fused': fused} if 'MulFrames' in cls else {}
    if not extra and not fused:
        pytest.skip('single-frame classes have one formulation')
    mod = getattr(pavenet_b200, cls)(dropout=0.0, **cfg, **extra).eval()
    mod.load_state_dict(state)
    lsi = O.level_start_index(inp['spatial_shapes'])
    out = mod(inp['query'], None, inp.get('value'), query_pos=inp.get('query_pos'),
              key_padding_mask=inp.get('key_padding_mask'),
              reference_points=inp['reference_points'], spatial_shapes=inp['spatial_shapes'],
              level_start_index=lsi)
    assert out.shape == c['out'].shape
    assert rel_err(out, c['out']) < 1e-5


def test_fused_softmax_does_not_overflow_where_reference_does(monkeypatch):
    """exp(logit) without max-subtraction overflows for logits > ~88
    (transformer.py:1736-1741, flagged BUG by the authors); the joint softmax
    stays finite and agrees with the reference formulation below that."""
    _with_oracle_op(monkeypatch)
    torch.manual_seed(0)
    kw = dict(embed_dims=32, num_heads=2, num_levels=2, num_points=3, dropout=0.0)
    fused = pavenet_b200.MulFramesMultiScaleDeformablePoseAttentionNumFrames3(fused=True, **kw).eval()
    plain = pavenet_b200.MulFramesMultiScaleDeformablePoseAttentionNumFrames3(fused=False, **kw).eval()
    with torch.no_grad():
        for p in fused.parameters():
            p.copy_(torch.randn_like(p) * 0.3)
        fused.attention_weights.bias.add_(100.0)
    plain.load_state_dict(fused.state_dict())
    shapes = torch.tensor([[4, 5], [2, 3]])
    lsi = O.level_start_index(shapes)
    args = dict(query_pos=None, reference_points=torch.rand(1, 3 * 4, 2, 6),
                spatial_shapes=shapes, level_start_index=lsi)
    q, v = torch.randn(4, 1, 32), torch.randn(26, 3, 32)
    assert torch.isfinite(fused(q, None, v, **args)).all()
    assert not torch.isfinite(plain(q, None, v, **args)).all()


def test_value_error_on_bad_reference_points(monkeypatch):
    _with_oracle_op(monkeypatch)
    shapes = torch.tensor([[2, 2]])
    lsi = torch.tensor([0])
    m = pavenet_b200.MultiScaleDeformableAttention(embed_dims=32, num_heads=1, num_levels=1)
    with pytest.raises(ValueError, match='2 or 4'):
        m(torch.randn(3, 1, 32), value=torch.randn(4, 1, 32),
          reference_points=torch.rand(1, 3, 1, 3), spatial_shapes=shapes, level_start_index=lsi)
    p = pavenet_b200.MultiScaleDeformablePoseAttention(embed_dims=32, num_heads=1, num_levels=1,
                                                       num_points=5)
    with pytest.raises(ValueError, match='2K'):
        p(torch.randn(3, 1, 32), None, torch.randn(4, 1, 32),
          reference_points=torch.rand(1, 3, 1, 8), spatial_shapes=shapes, level_start_index=lsi)


def test_clip_model_frozen_bn_fold_is_the_same_function():
    """The PAVE-Net step folds the backbone's frozen BatchNorms into the convolutions in front of
    them (clip_model.PaveNetR50.fold_frozen_bn); same features and same weight gradients as
    torchvision's conv -> bn -> relu path, checked in float64 on the CPU (pure PyTorch code)."""
    from pavenet_b200 import clip_model
    torch.manual_seed(0)
    m = clip_model.PaveNetR50(num_query=20).double().train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.5)
            mod.running_var.uniform_(0.5, 2.0)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.3)
    m.train()                                   # clears the fold cache
    x = torch.randn(1, 3, 3, 64, 96, dtype=torch.float64)
    names = ['layer4.2.conv3.weight', 'layer3.0.conv2.weight', 'layer2.0.downsample.0.weight']
    params = dict(m.named_parameters())
    res = {}
    for fold in (True, False):
        m.fold_frozen_bn = fold
        feats = m.extract_feat(x)
        loss = sum(t.square().sum() for t in feats)
        res[fold] = (feats, torch.autograd.grad(loss, [params[n] for n in names]))
    for a, b in zip(res[True][0], res[False][0]):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-10
    for a, b in zip(res[True][1], res[False][1]):
        assert float((a - b).abs().max() / b.abs().max()) < 1e-6


def test_bench_algorithmic_bytes_match_the_survey_figures():
    """bench.py's roofline numerator: config 2 forward 238.9 MB, backward
    409.6 MB, 9.73 KB per query; config 3 (T=5, P=17) 371.3 MB fp32 / 200.7 MB... per fwd+bwd."""
    import bench
    S = sum(h * w for h, w in bench.R50_LEVELS)
    assert S == 22223
    cfg2 = bench.algorithmic_bytes(dict(B=3, S=S, M=8, D=32, L=4, Q=S, P=4))
    assert round(cfg2['fwd'] / 1e6, 1) == 238.9 and round(cfg2['bwd'] / 1e6, 1) == 409.6
    assert round((cfg2['fwd'] + cfg2['bwd']) / (3 * S) / 1e3, 2) == 9.73
    cfg3 = bench.algorithmic_bytes(dict(B=1, S=5 * S, M=8, D=32, L=20, Q=300, P=17))
    assert round((cfg3['fwd'] + cfg3['bwd']) / 1e6, 1) == 371.3
    cfg1 = bench.algorithmic_bytes(dict(B=1, S=S, M=8, D=32, L=4, Q=300, P=17))
    assert round(cfg1['fwd'] / 1e6, 1) == 25.0 and round(cfg1['bwd'] / 1e6, 1) == 49.7


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to ours) runs without a GPU
    and prints ONE JSON line with the keys of the contract."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    proc = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference',
                           '--steps', '1', '--warmup', '3', '--workload', 'petr_cfg1'],
                          capture_output=True, text=True, timeout=600, cwd=root)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'queries/s' and d['higher_is_better'] is True
    assert d['metric'] == 'deform-attn fwd+bwd queries/s' and d['value'] > 0
    assert d['config']['workload'] == 'petr_cfg1' and d['config']['dims']['Q'] == 300
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'queries/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0


def test_integration_md_ctypes_stub_matches_the_abi():
    """INTEGRATION.md section 2b shows the binding a maintainer would paste into mmcv.  Execute that
    very text against the built library (no kernel is launched) and check that it declares the two
    entry points exactly as pavenet_b200._capi does."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, 'INTEGRATION.md')).read()
    sec = text[text.index('### 2b.'):text.index('### 2c.')]
    blocks = re.findall(r'```python\n(.*?)```', sec, flags=re.S)
    stub = [b for b in blocks if 'ctypes.CDLL' in b]
    assert len(stub) == 1
    code = stub[0].replace("ctypes.CDLL('libpavenet_msda.so')",
                           'ctypes.CDLL(%r)' % pavenet_b200._build.LIB_PATH)
    ns = {}
    exec(compile(code, 'INTEGRATION.md:2b', 'exec'), ns)
    lib = pavenet_b200._capi.load()
    assert len(ns['_lib'].msda_forward.argtypes) == len(lib.msda_forward.argtypes) == 16
    assert len(ns['_lib'].msda_backward.argtypes) == len(lib.msda_backward.argtypes) == 20
    for mine, theirs in ((ns['_lib'].msda_forward.argtypes, lib.msda_forward.argtypes),
                         (ns['_lib'].msda_backward.argtypes, lib.msda_backward.argtypes)):
        assert [ctypes.sizeof(a) for a in mine] == [ctypes.sizeof(a) for a in theirs]
    assert callable(ns['ext_module'].ms_deform_attn_forward)
    assert callable(ns['ext_module'].ms_deform_attn_backward)
    assert ns['_DT'] == {torch.float32: pavenet_b200._capi.MSDA_F32, torch.float64: pavenet_b200._capi.MSDA_F64,
                         torch.bfloat16: pavenet_b200._capi.MSDA_BF16}


def test_install_patches_a_live_mmcv_and_opera_tree(tmp_path, monkeypatch):
    """`pavenet_b200.install()` against a stub of the reference's package layout: the five things
    a maintainer needs swapped are swapped — `mmcv.ops.multi_scale_deform_attn.ext_module`
    and `.MultiScaleDeformableAttnFunction` (what third_party/mmcv/mmcv/utils/ext_loader.py:12-16
    and multi_scale_deform_attn.py:16-17,20 provide), the Function name imported into
    `opera.models.utils.transformer` (transformer.py:14), and the class names in mmcv's
    ATTENTION / FEEDFORWARD_NETWORK registries and opera's ATTENTION registry
    (mmcv/cnn/bricks/registry.py, opera/models/utils/builder.py:11-17)."""
    import importlib
    import sys
    import textwrap
    registry_src = textwrap.dedent('''
        class Registry(object):
            def __init__(self, name):
                self.name, self.module_dict = name, {}
            def register_module(self, name=None, force=False, module=None):
                if not force and name in self.module_dict:
                    raise KeyError(name)
                self.module_dict[name] = module
                return module
            def get(self, key):
                return self.module_dict.get(key)
    ''')
    files = {
        'mmcv/__init__.py': '',
        'mmcv/ops/__init__.py': '',
        'mmcv/ops/multi_scale_deform_attn.py':
            'ext_module = "ORIGINAL_EXT"\nclass MultiScaleDeformableAttnFunction(object):\n    pass\n',
        'mmcv/cnn/__init__.py': '',
        'mmcv/cnn/bricks/__init__.py': '',
        'mmcv/cnn/bricks/registry.py': registry_src + 'ATTENTION = Registry("attention")\n'
                                       'FEEDFORWARD_NETWORK = Registry("feed-forward Network")\n'
                                       'ATTENTION.register_module(name="MultiScaleDeformableAttention", module=object)\n',
        'opera/__init__.py': '',
        'opera/models/__init__.py': '',
        'opera/models/utils/__init__.py': '',
        'opera/models/utils/transformer.py':
            'from mmcv.ops.multi_scale_deform_attn import MultiScaleDeformableAttnFunction\n',
        'opera/models/utils/builder.py': registry_src + 'ATTENTION = Registry("attention")\n',
    }
    for rel, src in files.items():
        path = tmp_path / rel
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_text(src)
    monkeypatch.syspath_prepend(str(tmp_path))
    for name in [m for m in sys.modules if m == 'mmcv' or m.startswith('mmcv.')
                 or m == 'opera' or m.startswith('opera.')]:
        monkeypatch.delitem(sys.modules, name)
    import pavenet_b200
    from pavenet_b200 import functional, modules
    try:
        patched = pavenet_b200.install()
        msda = importlib.import_module('mmcv.ops.multi_scale_deform_attn')
        assert msda.ext_module is functional.ext_module
        assert msda.MultiScaleDeformableAttnFunction is functional.MultiScaleDeformableAttnFunction
        ot = importlib.import_module('opera.models.utils.transformer')
        assert ot.MultiScaleDeformableAttnFunction is functional.MultiScaleDeformableAttnFunction
        reg = importlib.import_module('mmcv.cnn.bricks.registry')
        assert reg.ATTENTION.get('MultiScaleDeformableAttention') is modules.MultiScaleDeformableAttention
        for cls in modules.MMCV_SCOPE_CLASSES:
            assert reg.ATTENTION.get(cls.__name__) is cls
        assert reg.FEEDFORWARD_NETWORK.get('FFN') is modules.FFN
        opera_reg = importlib.import_module('opera.models.utils.builder').ATTENTION
        for cls in modules.OPERA_SCOPE_CLASSES:
            assert opera_reg.get(cls.__name__) is cls
        assert {'mmcv.ops.multi_scale_deform_attn', 'opera.models.utils.transformer',
                'mmcv.FFN'} <= set(patched)
        assert 'opera.MulFramesMultiScaleDeformablePoseAttentionNumFrames3' in patched
    finally:
        for name in [m for m in sys.modules if m == 'mmcv' or m.startswith('mmcv.')
                     or m == 'opera' or m.startswith('opera.')]:
            sys.modules.pop(name, None)
