"""Registry glue: how configs pick the attention modules up by `type=` string.

The reference resolves `type='MultiScaleDeformableAttention'`,
`'mmcv.MulFramesMultiScaleDeformableAttentionNumFrames3'`,
`'opera.MulFramesMultiScaleDeformablePoseAttentionNumFrames3'` … through
mmcv's `ATTENTION` registry and opera's child registry
(third_party/mmcv/mmcv/cnn/bricks/registry.py, opera/models/utils/builder.py:11-17).

When mmcv is importable the classes of `pavenet_b200.modules` are registered
into the real registries under the reference's names (`force=True`, replacing
the originals) — see `install()`.  When it is not (this image), a minimal
stand-in with the same `register_module` / `build` / `get` surface is used so
configs written for the reference still build.
"""

__all__ = ['ATTENTION', 'OPERA_ATTENTION', 'FEEDFORWARD_NETWORK', 'build_attention',
           'build_feedforward_network', 'install']


class _MiniRegistry(object):
    """Just enough of `mmcv.utils.Registry` for `build_from_cfg`-style use."""

    def __init__(self, name, scope, parent=None):
        self.name, self.scope, self.parent = name, scope, parent
        self._module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if not force and key in self._module_dict:
                raise KeyError('%s is already registered in %s' % (key, self.name))
            self._module_dict[key] = cls
            return cls
        return _register(module) if module is not None else _register

    def get(self, key):
        scope, _, real = key.rpartition('.')
        reg = self
        if scope and scope != self.scope:
            # 'mmcv.X' asked of the opera registry walks up to the parent
            reg = self.parent if (self.parent and self.parent.scope == scope) else None
            if reg is None:
                return None
        found = reg._module_dict.get(real)
        if found is None and not scope and self.parent is not None:
            found = self.parent._module_dict.get(real)
        return found

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or 'type' not in cfg:
            raise KeyError('cfg must be a dict with a "type" key, got %r' % (cfg,))
        args = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop('type')
        cls = self.get(obj_type) if isinstance(obj_type, str) else obj_type
        if cls is None:
            raise KeyError('%s is not in the %s registry' % (obj_type, self.name))
        return cls(**args)

    def __contains__(self, key):
        return self.get(key) is not None


ATTENTION = _MiniRegistry('attention', scope='mmcv')
OPERA_ATTENTION = _MiniRegistry('attention', scope='opera', parent=ATTENTION)
FEEDFORWARD_NETWORK = _MiniRegistry('feed-forward Network', scope='mmcv')


def build_attention(cfg, default_args=None):
    """`mmcv.cnn.bricks.transformer.build_attention` for the classes of this package."""
    return OPERA_ATTENTION.build(cfg, default_args)


def build_feedforward_network(cfg, default_args=None):
    """`mmcv.cnn.bricks.transformer.build_feedforward_network` for this package's FFN."""
    return FEEDFORWARD_NETWORK.build(cfg, default_args)


def install():
    """Swap the B200 op into a live mmcv / opera installation.

    1. `mmcv.ops.multi_scale_deform_attn.{ext_module, MultiScaleDeformableAttnFunction}`
       are replaced, so every one of the reference's 18 attention classes (which
       all call `MultiScaleDeformableAttnFunction.apply`) runs on the new kernels
       unchanged.
    2. The fused module classes of `pavenet_b200.modules` (attention classes and
       the transformer layers' `FFN`) are registered over the reference's registry names.

    Returns the list of things patched; raises ImportError if mmcv is absent.
    """
    import importlib
    from . import functional, modules
    patched = []
    msda = importlib.import_module('mmcv.ops.multi_scale_deform_attn')
    msda.ext_module = functional.ext_module
    msda.MultiScaleDeformableAttnFunction = functional.MultiScaleDeformableAttnFunction
    patched.append('mmcv.ops.multi_scale_deform_attn')
    try:
        ot = importlib.import_module('opera.models.utils.transformer')
        ot.MultiScaleDeformableAttnFunction = functional.MultiScaleDeformableAttnFunction
        patched.append('opera.models.utils.transformer')
    except ImportError:
        pass
    from mmcv.cnn.bricks.registry import ATTENTION as MMCV_ATTENTION
    for cls in modules.MMCV_SCOPE_CLASSES:
        MMCV_ATTENTION.register_module(name=cls.__name__, force=True, module=cls)
        patched.append('mmcv.' + cls.__name__)
    from mmcv.cnn.bricks.registry import FEEDFORWARD_NETWORK as MMCV_FFN
    MMCV_FFN.register_module(name='FFN', force=True, module=modules.FFN)
    patched.append('mmcv.FFN')
    try:
        from opera.models.utils.builder import ATTENTION as OPERA_REG
        for cls in modules.OPERA_SCOPE_CLASSES:
            OPERA_REG.register_module(name=cls.__name__, force=True, module=cls)
            patched.append('opera.' + cls.__name__)
    except ImportError:
        pass
    return patched
