"""Seeded parameters and inputs of the production-geometry module fixtures
(`module_golden_256.npz`): embed_dims 256, 8 heads x 32 channels, 4 levels, P = 4 / 15 / 17
(configs/videopose/2025-2-13/2025_2_13_res50_num_frames_3_posetrack17.py:53-107,
configs/petr/petr_r50_16x2_100e_coco.py).

Used by BOTH `gen_golden.py` (which executes the reference classes on these tensors in the
build container) and the tests (which run this repository's modules / the CPU composition
oracle on the same tensors).  Storing the tensors instead would cost ~20 MB of state dicts;
`checksum()` guards against the two sides drifting apart.
"""
import torch

SHAPES = [(6, 9), (3, 5), (2, 3), (1, 2)]     # 4 levels, S = 77
C, M, L = 256, 8, 4

CASES = [
    dict(name='encoder', where='msda', cls='MultiScaleDeformableAttention',
         cfg=dict(embed_dims=C, num_heads=M, num_levels=L, num_points=4),
         kind='encoder', B=2, param_seed=101, input_seed=201),
    dict(name='pose', where='ot', cls='MultiScaleDeformablePoseAttention',
         cfg=dict(embed_dims=C, num_heads=M, num_levels=L, num_points=17),
         kind='pose', B=2, Q=9, param_seed=102, input_seed=202),
    dict(name='mf_pose3', where='ot', cls='MulFramesMultiScaleDeformablePoseAttentionNumFrames3',
         cfg=dict(embed_dims=C, num_heads=M, num_levels=L, num_points=15),
         kind='mf_pose', T=3, B=2, Q=7, param_seed=103, input_seed=203),
    dict(name='mf_pose5', where='ot', cls='MulFramesMultiScaleDeformablePoseAttentionNumFrames5',
         cfg=dict(embed_dims=C, num_heads=M, num_levels=L, num_points=17),
         kind='mf_pose', T=5, B=1, Q=6, param_seed=104, input_seed=204),
    dict(name='mf_joint3', where='msda', cls='MulFramesMultiScaleDeformableAttentionNumFrames3',
         cfg=dict(embed_dims=C, num_heads=M, num_levels=L, num_points=4),
         kind='mf_joint', T=3, B=3, Q=15, cuda_like=True, param_seed=105, input_seed=205),
    dict(name='mf_joint5', where='msda', cls='MulFramesMultiScaleDeformableAttentionNumFrames5',
         cfg=dict(embed_dims=C, num_heads=M, num_levels=L, num_points=4),
         kind='mf_joint', T=5, B=2, Q=17, cuda_like=True, param_seed=106, input_seed=206),
]


def case(name):
    for c in CASES:
        if c['name'] == name:
            return c
    raise KeyError(name)


def randomise(module, seed):
    """Every parameter ~ N(0, s^2), drawn in sorted-name order from one CPU generator.

    Each parameter first gets storage of its own: the reference's multi-frame `init_weights`
    assigns ONE `grid_init` tensor to the `.data` of all per-frame `sampling_offsets` biases
    (multi_scale_deform_attn.py:1374-1376), so in a freshly built reference module those
    biases alias each other and an in-place fill of one would overwrite the others."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            p.data = p.data.clone()
            if 'sampling_offsets' in name:
                scale = 0.05
            elif p.dim() == 2:
                scale = 0.06            # ~ 1/sqrt(256): activations stay O(1)
            else:
                scale = 0.3
            p.copy_((torch.randn(p.shape, generator=g) * scale).to(p.device))


def make_inputs(c):
    g = torch.Generator().manual_seed(c['input_seed'])
    shapes = torch.as_tensor(SHAPES, dtype=torch.long)
    S = int(shapes.prod(1).sum())

    def rnd(*shape):
        return torch.randn(*shape, generator=g)

    def mask(*shape):
        return torch.rand(*shape, generator=g) < 0.15

    kind, B = c['kind'], c['B']
    P = c['cfg']['num_points']
    if kind == 'encoder':
        inp = dict(query=rnd(S, B, C), query_pos=rnd(S, B, C), key_padding_mask=mask(B, S),
                   reference_points=torch.rand(B, S, L, 2, generator=g))
    elif kind == 'pose':
        Q = c['Q']
        inp = dict(query=rnd(Q, B, C), query_pos=rnd(Q, B, C), value=rnd(S, B, C),
                   key_padding_mask=mask(B, S),
                   reference_points=torch.rand(B, Q, L, 2 * P, generator=g))
    elif kind == 'mf_pose':
        Q, T = c['Q'], c['T']
        inp = dict(query=rnd(Q, B, C), query_pos=rnd(Q, B, C), value=rnd(S, B * T, C),
                   key_padding_mask=mask(B * T, S),
                   reference_points=torch.rand(B, T * Q, L, 2 * P, generator=g))
    elif kind == 'mf_joint':
        Q, T = c['Q'], c['T']
        inp = dict(query=rnd(Q, B, C), query_pos=rnd(Q, B, C), value=rnd(S, B, T, C),
                   key_padding_mask=mask(B, T, S),
                   reference_points=torch.rand(T * B, Q, L, 2, generator=g))
    else:
        raise ValueError(kind)
    inp['spatial_shapes'] = shapes
    return inp


def grad_output(c, shape):
    g = torch.Generator().manual_seed(c['input_seed'] + 5000)
    return torch.randn(*shape, generator=g)


def projection_vectors(c, pname, shape):
    """(u, v): seeded unit-variance vectors for u^T dW (length in) and dW v (length out)."""
    seed = c['param_seed'] * 1000 + sum(ord(ch) for ch in pname)
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape[0], generator=g), torch.randn(shape[1], generator=g)


def checksum(module, inp):
    """A few moments of the regenerated tensors, compared loosely (1e-5) by the tests."""
    vals = []
    for _, p in sorted(module.named_parameters()):
        p = p.detach().double().cpu()
        vals += [p.sum(), p.abs().sum()]
    for k in sorted(inp):
        t = inp[k].detach().double().cpu()
        vals += [t.sum(), t.abs().sum()]
    return torch.stack(vals)
