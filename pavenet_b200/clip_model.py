"""A self-contained restatement of the PAVE-Net R-50 clip training step
 around the B200 attention modules.

Purpose: measure clips/s of a clip-sharded training step at 1/2/4/8 GPUs
(BASELINE.json config 4).  The reference model cannot be imported in this
environment (mmcv / mmdet / opera dependencies are absent, SURVEY.md
section 8c), so this is a from-scratch model that follows the reference's
graph — tensor shapes, layer counts and op sequence — for the canonical T=3
configuration `configs/videopose/2025-2-13/2025_2_13_res50_num_frames_3_posetrack17.py:8-153`:

  backbone   ResNet-50, frames folded into the batch, stage 1 + all BN frozen
             (mmdet/models/backbones/resnet.py:632-653)
  neck       ChannelMapper: 1x1 conv + GN(32) on C3..C5, 3x3/2 conv + GN for the 4th level
             (mmdet/models/necks/channel_mapper.py:53-79)
  encoder    6 x [MultiScaleDeformableAttention, LN, FFN(1024), LN] over all
             T frames (opera/models/utils/transformer.py:21279-21330)
  two-stage  proposals from the current frame's memory, top-300 (transformer.py:21340-21403)
  pose dec.  3 x [MHA, LN, MulFramesMultiScaleDeformablePoseAttentionNumFrames3, LN, FFN, LN]
             with per-frame keypoint refinement, references NOT detached
             (transformer.py:6711-6746)
  losses     Hungarian matching (focal + L1 + OKS cost), focal cls loss,
             RLE keypoint loss with a RealNVP prior
             (opera/models/dense_heads/videopose_head_mul_frames.py:795-1010; losses/oks_loss.py:162-195)
  joint dec. 2 x [MHA over the K keypoint queries, LN,
             MulFramesMultiScaleDeformableAttentionNumFrames3, LN, FFN, LN] per matched person
             (videopose_head_mul_frames.py:569-742; transformer.py:21458-21536;
             mmdet/models/utils/transformer.py:843-886)
  optimizer  AdamW lr 2e-5, wd 1e-4, 0.1x lr for backbone / sampling_offsets, grad-clip 0.1

Parity status of THIS file: unpinned (nothing to run it against here); the
attention modules inside it are the pinned ones.  It is a throughput vehicle,
not a re-implementation of the reference's training recipe.
"""
import math

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import modules
from .modules import (MulFramesMultiScaleDeformableAttentionNumFrames3,
                      MulFramesMultiScaleDeformablePoseAttentionNumFrames3,
                      MultiScaleDeformableAttention)

POSETRACK_SIGMAS = torch.tensor([.26, .79, .79, .72, .62, .79, .72, .62, 1.07, .87, .89, 1.07, .87,
                                 .89, .25][:15]) / 10.0  # nose, head, neck(ish), shoulders ... ankles


def inverse_sigmoid(x, eps=1e-5):
    x = x.clamp(0, 1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def reduce_mean(t):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = t.clone()
        dist.all_reduce(t.div_(dist.get_world_size()))
    return t


class SinePositionalEncoding(nn.Module):
    """mmcv SinePositionalEncoding(num_feats=128, normalize=True, offset=-0.5)."""

    def __init__(self, num_feats=128, temperature=10000, scale=2 * math.pi, offset=-0.5):
        super().__init__()
        self.num_feats, self.temperature, self.scale, self.offset = num_feats, temperature, scale, offset

    def forward(self, mask):                      # (N, H, W) bool, True = padding
        not_mask = (~mask).float()
        y = not_mask.cumsum(1)
        x = not_mask.cumsum(2)
        y = (y + self.offset) / (y[:, -1:, :] + 1e-6) * self.scale
        x = (x + self.offset) / (x[:, :, -1:] + 1e-6) * self.scale
        dim_t = torch.arange(self.num_feats, dtype=torch.float32, device=mask.device)
        dim_t = self.temperature ** (2 * (dim_t // 2) / self.num_feats)
        px, py = x[..., None] / dim_t, y[..., None] / dim_t
        px = torch.stack((px[..., 0::2].sin(), px[..., 1::2].cos()), 4).flatten(3)
        py = torch.stack((py[..., 0::2].sin(), py[..., 1::2].cos()), 4).flatten(3)
        return torch.cat((py, px), 3).permute(0, 3, 1, 2)


def FFN(dims=256, hidden=1024, drop=0.1):
    """feedforward_channels=1024, ffn_dropout=0.1 (config lines 54-56): x + drop(fc2(drop(relu(fc1(x)))))."""
    return modules.FFN(embed_dims=dims, feedforward_channels=hidden, ffn_drop=drop)


class SelfAttention(nn.Module):
    """mmcv MultiheadAttention wrapper: q = k = x + pos, v = x, residual + dropout."""

    def __init__(self, dims=256, heads=8, drop=0.1):
        super().__init__()
        self.attn = nn.MultiheadAttention(dims, heads, dropout=drop)
        self.drop = nn.Dropout(drop)

    def forward(self, x, pos):
        q = x + pos
        return x + self.drop(self.attn(q, q, x, need_weights=False)[0])


class EncoderLayer(nn.Module):
    def __init__(self, value_dtype):
        super().__init__()
        # batch_first: tokens stay (frames, S, 256) through the encoder, so no projection needs a
        # transposing copy (a constructor option of the reference class, multi_scale_deform_attn.py:244)
        self.attn = MultiScaleDeformableAttention(embed_dims=256, value_dtype=value_dtype, batch_first=True)
        self.norm1, self.ffn, self.norm2 = modules.LayerNorm(256), FFN(), modules.LayerNorm(256)

    def forward(self, x, pos, mask, ref, shapes, lsi):
        x = self.norm1(self.attn(x, query_pos=pos, key_padding_mask=mask, reference_points=ref,
                                 spatial_shapes=shapes, level_start_index=lsi))
        return self.norm2(self.ffn(x))


class DecoderLayer(nn.Module):
    def __init__(self, cross):
        super().__init__()
        self.self_attn, self.cross = SelfAttention(), cross
        self.norm1, self.norm2, self.ffn, self.norm3 = (modules.LayerNorm(256), modules.LayerNorm(256), FFN(),
                                                        modules.LayerNorm(256))

    def forward(self, x, pos, **cross_kw):
        x = self.norm1(self.self_attn(x, pos))
        x = self.norm2(self.cross(x, None, query_pos=pos, **cross_kw))
        return self.norm3(self.ffn(x))


def mlp(inp, hidden, out, n_hidden=2):
    layers, d = [], inp
    for _ in range(n_hidden):
        layers += [nn.Linear(d, hidden), nn.ReLU()]
        d = hidden
    return nn.Sequential(*layers, nn.Linear(d, out))


class RealNVP(nn.Module):
    """2-D flow prior of the RLE loss (3 coupling pairs, 64 hidden units)."""

    def __init__(self, n=6, hidden=64):
        super().__init__()
        self.register_buffer('masks', torch.tensor([[0., 1.], [1., 0.]] * (n // 2)))
        self.s = nn.ModuleList(nn.Sequential(nn.Linear(2, hidden), nn.LeakyReLU(), nn.Linear(hidden, hidden),
                                             nn.LeakyReLU(), nn.Linear(hidden, 2), nn.Tanh()) for _ in range(n))
        self.t = nn.ModuleList(nn.Sequential(nn.Linear(2, hidden), nn.LeakyReLU(), nn.Linear(hidden, hidden),
                                             nn.LeakyReLU(), nn.Linear(hidden, 2)) for _ in range(n))

    def log_prob(self, x):
        log_det, z = x.new_zeros(x.shape[0]), x
        for i in reversed(range(len(self.s))):
            m = self.masks[i]
            z_ = m * z
            s, t = self.s[i](z_) * (1 - m), self.t[i](z_) * (1 - m)
            z = (1 - m) * (z - t) * torch.exp(-s) + z_
            log_det = log_det - s.sum(1)
        return -0.5 * (z ** 2).sum(1) - math.log(2 * math.pi) + log_det


class PaveNetR50(nn.Module):
    def __init__(self, num_frames=3, num_keypoints=15, num_query=300, value_dtype=None):
        super().__init__()
        import torchvision
        if num_frames != 3:
            raise ValueError('this restatement covers the canonical T=3 configuration')
        self.T, self.K, self.Qn = num_frames, num_keypoints, num_query
        r = torchvision.models.resnet50(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layer1, self.layer2, self.layer3, self.layer4 = r.layer1, r.layer2, r.layer3, r.layer4
        for m in self.modules():                      # norm_eval=True, BN requires_grad=False
            if isinstance(m, nn.BatchNorm2d):
                for p in m.parameters():
                    p.requires_grad_(False)
        for p in list(self.stem.parameters()) + list(self.layer1.parameters()):   # frozen_stages=1
            p.requires_grad_(False)
        self.lateral = nn.ModuleList(nn.Sequential(nn.Conv2d(c, 256, 1), nn.GroupNorm(32, 256))
                                     for c in (512, 1024, 2048))
        self.extra = nn.Sequential(nn.Conv2d(2048, 256, 3, stride=2, padding=1), nn.GroupNorm(32, 256))
        self.pos_enc = SinePositionalEncoding()
        self.level_embeds = nn.Parameter(torch.randn(4, 256))
        self.encoder = nn.ModuleList(EncoderLayer(value_dtype) for _ in range(6))
        self.enc_output, self.enc_output_norm = nn.Linear(256, 256), modules.LayerNorm(256)
        K = num_keypoints
        self.decoder = nn.ModuleList(
            DecoderLayer(MulFramesMultiScaleDeformablePoseAttentionNumFrames3(
                num_points=K, embed_dims=256, value_dtype=value_dtype)) for _ in range(3))
        self.refine_decoder = nn.ModuleList(
            DecoderLayer(MulFramesMultiScaleDeformableAttentionNumFrames3(
                embed_dims=256, im2col_step=128, value_dtype=value_dtype)) for _ in range(2))
        self.query_embedding = nn.Embedding(num_query, 512)
        self.refine_query_embedding = nn.Embedding(K, 512)
        self.cls_branches = nn.ModuleList(nn.Linear(256, 1) for _ in range(4))
        self.kpt_branches = nn.ModuleList(mlp(256, 512, 2 * K) for _ in range(4))
        self.pre_kpt_branches = nn.ModuleList(mlp(256, 512, 2 * K) for _ in range(3))
        self.next_kpt_branches = nn.ModuleList(mlp(256, 512, 2 * K) for _ in range(3))
        self.sigma_branches = nn.ModuleList(mlp(256, 256, 2 * K, n_hidden=1) for _ in range(4))
        self.refine_kpt = nn.ModuleList(nn.ModuleList(mlp(256, 256, 2) for _ in range(2)) for _ in range(3))
        self.refine_sigma = nn.ModuleList(mlp(256, 256, 2, n_hidden=1) for _ in range(2))
        self.flow = RealNVP()
        nn.init.constant_(self.cls_branches[0].bias, -4.6)
        self._geo_cache, self._graphed, self._fold_cache = {}, None, {}

    def train(self, mode=True):
        super().train(mode)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        self.stem.eval()
        self.layer1.eval()
        self._fold_cache = {}
        return self

    def load_state_dict(self, *args, **kwargs):
        self._fold_cache = {}
        return super().load_state_dict(*args, **kwargs)

    # ---- frozen BatchNorm folded into the convolution in front of it -------------------------
    #: Every BatchNorm of the backbone is frozen (norm_eval=True, requires_grad=False:
    #: mmdet/models/backbones/resnet.py:632-653), i.e. a per-channel affine y = s * conv(x) + t with
    #: constant s, t.  Folding it into the convolution (w' = s * w, bias t) is the same function and
    #: removes one read + write of every backbone activation in the forward and one in the backward
    #: (53 `bn_fw_inf` launches, 5.5 ms of a 60 ms step on B200).  False: torchvision's module-by-module path.
    fold_frozen_bn = True

    def _bn_affine(self, bn):
        hit = self._fold_cache.get(id(bn))
        if hit is None:
            with torch.no_grad():
                s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
                hit = (s.view(-1, 1, 1, 1).contiguous(), (bn.bias - bn.running_mean * s).contiguous())
            self._fold_cache[id(bn)] = hit
        return hit

    def _conv_bn(self, conv, bn, x):
        s, t = self._bn_affine(bn)
        if conv.weight.requires_grad:
            w = conv.weight * s
        else:                                     # frozen stage: fold once
            w = self._fold_cache.get(id(conv))
            if w is None:
                with torch.no_grad():
                    w = (conv.weight * s).contiguous(memory_format=torch.channels_last)
                self._fold_cache[id(conv)] = w
        return F.conv2d(x, w, t, conv.stride, conv.padding, conv.dilation, conv.groups)

    def _bottleneck(self, blk, x):
        """torchvision Bottleneck.forward with the BatchNorms folded."""
        out = F.relu_(self._conv_bn(blk.conv1, blk.bn1, x))
        out = F.relu_(self._conv_bn(blk.conv2, blk.bn2, out))
        out = self._conv_bn(blk.conv3, blk.bn3, out)
        idt = x if blk.downsample is None else self._conv_bn(blk.downsample[0], blk.downsample[1], x)
        return F.relu_(out.add_(idt))

    def _res_layer(self, layer, x):
        if not self.fold_frozen_bn:
            return layer(x)
        for blk in layer:
            x = self._bottleneck(blk, x)
        return x

    def _stem(self, x):
        if not self.fold_frozen_bn:
            return self.stem(x)
        conv, bn, _, pool = self.stem
        return pool(F.relu_(self._conv_bn(conv, bn, x)))

    #: diagnostic hook: set to a callable(name) to get a call at each phase boundary of
    #: forward_train (tools/phase_step.py synchronises and timestamps there)
    phase_hook = None

    def _mark(self, name):
        if self.phase_hook is not None:
            self.phase_hook(name)

    # ------------------------------------------------------------------ graph
    def extract_feat(self, images):                   # (Bc, T, 3, H, W) -> 4 levels of (Bc*T, 256, h, w)
        return list(self._stage_top(*self._stage_trunk(images)))

    # The backbone is two stages so that the gradient exchange can overlap with it: when the
    # backward leaves `top` (layer4 + neck, 71 % of the backbone's parameters) that bucket is
    # all-reduced underneath the backward of `trunk` (layer2-3 over the full-resolution maps).
    def _stage_trunk(self, images):                   # -> c3 (stride 8), c4 (stride 16)
        x = images.flatten(0, 1).contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            x = self._res_layer(self.layer1, self._stem(x))
        c3 = self._res_layer(self.layer2, x)
        c4 = self._res_layer(self.layer3, c3)
        return c3, c4

    def _stage_top(self, c3, c4):                     # -> the 4 levels of (Bc*T, 256, h, w)
        c5 = self._res_layer(self.layer4, c4)
        return tuple([l(c) for l, c in zip(self.lateral, (c3, c4, c5))] + [self.extra(c5)])

    @staticmethod
    def reference_grid(shapes_list, device):
        pts = []
        for h, w in shapes_list:
            ys = (torch.arange(h, device=device, dtype=torch.float32) + 0.5) / h
            xs = (torch.arange(w, device=device, dtype=torch.float32) + 0.5) / w
            yy, xx = torch.meshgrid(ys, xs, indexing='ij')
            pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
        return torch.cat(pts)                          # (S, 2)

    # ---- constants of one input geometry (cached: nothing here depends on the data) ----
    def _geometry(self, images):
        Bc, T = images.shape[:2]
        H, W = images.shape[-2:]
        key = (Bc, T, H, W, images.device)
        geo = self._geo_cache.get(key)
        if geo is not None:
            return geo
        dev = images.device

        def down(n, times):
            for _ in range(times):
                n = (n - 1) // 2 + 1
            return n
        shapes_list = [(down(H, k), down(W, k)) for k in (3, 4, 5, 6)]      # strides 8 / 16 / 32 / 64
        shapes = torch.tensor(shapes_list, device=dev)
        lsi = torch.cat([shapes.new_zeros(1), shapes.prod(1).cumsum(0)[:-1]])
        masks = [torch.zeros((Bc * T, h, w), dtype=torch.bool, device=dev) for h, w in shapes_list]
        with torch.no_grad():
            pos_sine = [self.pos_enc(m) for m in masks]                       # (Bc*T, 256, h, w) each
        mask_flat = torch.cat([m.flatten(1) for m in masks], 1)               # (Bc*T, S)
        S = mask_flat.shape[1]
        grid = self.reference_grid(shapes_list, dev)
        geo = dict(shapes_list=shapes_list, shapes=shapes, lsi=lsi, pos_sine=pos_sine, mask_flat=mask_flat,
                   grid=grid, ref_enc=grid[None, :, None, :].expand(Bc * T, S, 4, 2).contiguous(),  # valid_ratios == 1
                   proposals=inverse_sigmoid(grid)[None].expand(Bc, S, 2).contiguous(),
                   wh=images.new_tensor([W, H]))
        self._geo_cache[key] = geo
        return geo

    # ---- the three static-shape stages (eager, or one CUDA graph each: enable_graphs) ----
    def _stage_encoder(self, f0, f1, f2, f3, p0, p1, p2, p3, mask_flat, ref_enc, shapes, lsi):
        feats, pos_sine = (f0, f1, f2, f3), (p0, p1, p2, p3)
        pos = torch.cat([(p + self.level_embeds[i].view(1, -1, 1, 1)).flatten(2)
                         for i, p in enumerate(pos_sine)], 2).transpose(1, 2).contiguous()  # (Bc*T, S, 256)
        x = torch.cat([f.flatten(2) for f in feats], 2).transpose(1, 2).contiguous()     # (Bc*T, S, 256)
        for layer in self.encoder:
            x = layer(x, pos, mask_flat, ref_enc, shapes, lsi)
        return x

    def _stage_decoder(self, x, proposals, mask_flat, shapes, lsi):
        """two-stage proposals from the current frame + the pose decoder.
        x (Bc*T, S, 256) -> enc_cls, enc_kpt, enc_sigma, then (cls, kpt, sigma) per decoder layer."""
        T, K = self.T, self.K
        Bc, S = x.shape[0] // T, x.shape[1]
        memory = x.transpose(0, 1)             # (S, Bc*T, 256) view: what the decoders' modules take
        now = x[T // 2::T]                                                              # (Bc, S, 256)
        out_mem = self.enc_output_norm(self.enc_output(now))
        enc_cls = self.cls_branches[3](out_mem)
        enc_kpt = self.kpt_branches[3](out_mem).view(Bc, S, K, 2) + proposals[:, :, None, :]
        enc_kpt = enc_kpt.flatten(2)
        enc_sigma = self.sigma_branches[3](out_mem).sigmoid()
        topk = enc_cls[..., 0].topk(self.Qn, dim=1)[1]
        ref = torch.gather(enc_kpt, 1, topk[..., None].expand(-1, -1, 2 * K)).detach().sigmoid()
        tgt = torch.gather(out_mem, 1, topk[..., None].expand(-1, -1, 256)).detach()
        ref = ref.repeat(1, T, 1)                                                       # (Bc, T*Q, 2K)
        q_pos, q = self.query_embedding.weight.split(256, 1)
        query = (tgt + q[None]).permute(1, 0, 2)                                        # (Q, Bc, 256)
        q_pos = q_pos[None].expand(Bc, -1, -1).permute(1, 0, 2)
        outs = [enc_cls, enc_kpt, enc_sigma]
        for lid, layer in enumerate(self.decoder):
            ref_in = ref[:, :, None, :].expand(-1, -1, 4, -1)                           # (Bc, T*Q, L, 2K)
            query = layer(query, q_pos, value=memory, key_padding_mask=mask_flat, reference_points=ref_in,
                          spatial_shapes=shapes, level_start_index=lsi)
            o = query.permute(1, 0, 2)
            deltas = torch.cat([self.pre_kpt_branches[lid](o), self.kpt_branches[lid](o),
                                self.next_kpt_branches[lid](o)], 1)
            ref = (deltas + inverse_sigmoid(ref)).sigmoid()                             # not detached
            outs += [self.cls_branches[lid](o), ref.view(Bc, T, self.Qn, 2 * K),
                     self.sigma_branches[lid](o).sigmoid()]
        return tuple(outs)

    def enable_graphs(self, enabled=True):
        """Run the three static-shape stages (backbone + neck, encoder, two-stage + pose decoder) as
        CUDA graphs, forward and backward (`graphs.GraphedStage`): the step issues ~6 300 kernels,
        and on one B200 the host cannot launch them as fast as the GPU retires them.  The joint
        decoder's shapes follow the number of matched persons, so it keeps one graph per count;
        matching and the losses (host-side Hungarian assignment) stay eager."""
        from . import graphs
        if not enabled:
            self._graphed = None
            return self
        mods_trunk = [self.stem, self.layer1, self.layer2, self.layer3]
        mods_top = [self.layer4, self.lateral, self.extra]
        mods_decoder = [self.enc_output, self.enc_output_norm, self.cls_branches, self.kpt_branches,
                        self.sigma_branches, self.pre_kpt_branches, self.next_kpt_branches, self.decoder,
                        self.query_embedding]
        self._graphed = dict(
            trunk=graphs.GraphedStage(self._stage_trunk, mods_trunk),
            top=graphs.GraphedStage(self._stage_top, mods_top),
            encoder=graphs.GraphedStage(self._stage_encoder, [self.encoder], [self.level_embeds]),
            decoder=graphs.GraphedStage(self._stage_decoder, mods_decoder),
            # keyed on the per-clip matched-person counts (data dependent): capture a count pattern
            # only when it comes back, keep a handful of graphs, run eagerly otherwise
            joint=graphs.GraphedStage(self._stage_joint, [self.refine_decoder, self.refine_kpt, self.refine_sigma,
                                                          self.refine_query_embedding],
                                      max_graphs=6, capture_after=2))
        return self

    #: set by FlatGradients(model, overlap=True): gradient buckets are all-reduced from autograd
    #: hooks as the backward pass completes them
    grad_exchange = None

    def gradient_buckets(self):
        """Trainable parameters in the order the backward pass completes their gradients:
        [everything after the encoder (pose decoder, heads, joint decoder) | encoder |
         backbone top (layer4 + neck) | backbone trunk (layer2-3)]."""
        in_trunk = {id(p) for m in (self.stem, self.layer1, self.layer2, self.layer3) for p in m.parameters()}
        in_top = {id(p) for m in (self.layer4, self.lateral, self.extra) for p in m.parameters()}
        in_encoder = {id(p) for p in self.encoder.parameters()} | {id(self.level_embeds)}
        late, enc, top, trunk = [], [], [], []
        for p in self.parameters():
            if not p.requires_grad:
                continue
            (trunk if id(p) in in_trunk else top if id(p) in in_top else enc if id(p) in in_encoder
             else late).append(p)
        return [late, enc, top, trunk]

    def _exchange_hook(self, tensor, n_buckets):
        """When the gradient of `tensor` (a stage boundary) is ready, the stages behind it have
        finished their backward: their gradient buckets can go out."""
        ex = self.grad_exchange
        if ex is not None and tensor.requires_grad and torch.is_grad_enabled():
            def hook(grad, ex=ex, n=n_buckets):
                ex.launch_through(n)
                return None
            tensor.register_hook(hook)

    def library_launches_replayed(self):
        """Kernels of libpavenet_msda.so launched through CUDA-graph replays so far (the
        library's own counter only sees launches made outside graphs)."""
        graphed = getattr(self, '_graphed', None) or {}
        return sum(st.stats['library_launches_replayed'] for st in graphed.values())

    def _run_stage(self, name, fn, *args):
        graphed = getattr(self, '_graphed', None)
        if graphed is not None and self.training and torch.is_grad_enabled():
            return graphed[name](*args)
        return fn(*args)

    def forward_train(self, images, gt_kpts, gt_areas):
        """images (Bc, T, 3, H, W); gt_kpts: per clip (G_i, T, K, 3) normalised x, y, visibility;
        gt_areas: per clip (G_i,) normalised box areas.  Returns a dict of losses."""
        T = self.T
        geo = self._geometry(images)
        shapes, lsi, mask_flat = geo['shapes'], geo['lsi'], geo['mask_flat']
        c3, c4 = self._run_stage('trunk', self._stage_trunk, images)
        self._exchange_hook(c4, 3)             # backward reaches the trunk: decoders, encoder, top done
        feats = self._run_stage('top', self._stage_top, c3, c4)
        if [tuple(f.shape[-2:]) for f in feats] != geo['shapes_list']:
            raise RuntimeError('backbone produced %r, expected %r'
                               % ([tuple(f.shape[-2:]) for f in feats], geo['shapes_list']))
        self._mark('backbone+neck')
        self._exchange_hook(feats[0], 2)       # backward reaches the backbone: decoders + encoder done
        x = self._run_stage('encoder', self._stage_encoder, *feats, *geo['pos_sine'], mask_flat,
                            geo['ref_enc'], shapes, lsi)
        self._exchange_hook(x, 1)              # backward reaches the encoder: decoders done
        memory = x.transpose(0, 1)             # (S, Bc*T, 256) view: what the decoders' modules take
        self._mark('encoder')
        outs = self._run_stage('decoder', self._stage_decoder, x, geo['proposals'], mask_flat, shapes, lsi)
        enc_cls, enc_kpt, enc_sigma = outs[:3]
        cls_out, kpt_out, sigma_out = outs[3::3], outs[4::3], outs[5::3]

        self._mark('two-stage + pose decoder')
        wh = geo['wh']
        stages = [(enc_cls, enc_kpt.sigmoid()[:, None].expand(-1, T, -1, -1), enc_sigma, 'enc')] + \
                 [(c, k, s, 'd%d' % i) for i, (c, k, s) in enumerate(zip(cls_out, kpt_out, sigma_out))]
        # every stage's Hungarian assignment behind ONE device->host synchronisation
        matches = self.match_all(stages, gt_kpts, gt_areas, wh)
        losses, rle_terms = {}, []
        for (cls, kpt, sigma, tag), match in zip(stages, matches):
            losses[tag + '.loss_cls'] = self.cls_loss(cls, match)
            rle_terms.append((tag + '.loss_kpt',) + self.kpt_terms(kpt, sigma, match, gt_kpts))
        self._mark('matching + losses')
        rle_terms += self.refine(memory, mask_flat, shapes, lsi, kpt_out[-1], matches[-1], gt_kpts)
        self._mark('joint decoder')
        losses.update(self.rle_all(rle_terms))
        self._mark('keypoint losses')
        return losses

    # ------------------------------------------------------------------ losses
    def rle_all(self, terms):
        """All residual-log-likelihood keypoint losses of the step through ONE evaluation of the flow
        prior: terms = [(name, pred, sigma, target, weight)], each (N_i, K, 2); loss_i is the sum over
        its rows divided by its own number of valid coordinates (oks_loss.py:162-195)."""
        sizes = [t[1].shape[0] for t in terms]
        pred, sigma, target, weight = (torch.cat([t[i] for t in terms]) for i in (1, 2, 3, 4))
        n_valid = torch.stack([t[4].sum() for t in terms]).detach()
        n_valid = reduce_mean(n_valid).clamp(min=1)                                     # stays on the device
        bar_mu = (pred - target) / sigma
        log_phi = self.flow.log_prob(bar_mu.reshape(-1, 2)).reshape(pred.shape[0], -1, 1)
        nf = (torch.log(sigma) - log_phi) * weight[:, :, :1]
        amp = 1 / math.sqrt(2 * math.pi)
        logq = (torch.log(sigma / amp) + (target - pred).abs() / (math.sqrt(2) * sigma + 1e-9)) * weight
        per_row = (nf + logq).sum((1, 2))
        seg = torch.repeat_interleave(torch.arange(len(sizes), device=pred.device),
                                      torch.tensor(sizes, device=pred.device), output_size=sum(sizes))
        totals = torch.zeros(len(sizes), device=pred.device).index_add(0, seg, per_row) / n_valid
        return {t[0]: totals[i] for i, t in enumerate(terms)}

    @torch.no_grad()
    def match_cost(self, cls, kpt_now, gts, areas, wh):
        """Assignment cost of one clip and stage, (N, G): 2*focal + 70*L1 + 7*(1-OKS)."""
        p = cls.sigmoid()[:, 0]
        cost_cls = (0.25 * (1 - p) ** 2 * -(p + 1e-12).log() - 0.75 * p ** 2 * -(1 - p + 1e-12).log())
        pred = kpt_now.view(-1, 1, self.K, 2)
        tgt, vis = gts[None, :, :, :2], (gts[None, :, :, 2] > 0).float()
        l1 = ((pred - tgt).abs().sum(-1) * vis).sum(-1) / vis.sum(-1).clamp(min=1)
        d2 = (((pred - tgt) * wh) ** 2).sum(-1)
        var = (2 * POSETRACK_SIGMAS.to(pred.device)) ** 2
        oks = (torch.exp(-d2 / (2 * (areas[None, :, None] * wh.prod()).clamp(min=1) * var)) * vis).sum(-1)
        oks = oks / vis.sum(-1).clamp(min=1)
        return 2.0 * cost_cls[:, None] + 70.0 * l1 + 7.0 * (1 - oks)

    @torch.no_grad()
    def match_all(self, stages, gt_kpts, gt_areas, wh):
        """Hungarian assignment for every stage and clip: the cost matrices are computed on the GPU,
        copied to pinned host memory together, and solved after a single synchronisation.
        Returns, per stage, a list over clips of (rows, cols) index tensors."""
        from scipy.optimize import linear_sum_assignment
        now = self.T // 2
        dev = stages[0][0].device
        pending = []
        for cls, kpt, _, _ in stages:
            per_clip = []
            for b in range(cls.shape[0]):
                if gt_kpts[b].shape[0] == 0:
                    per_clip.append(None)
                    continue
                cost = self.match_cost(cls[b], kpt[b, now], gt_kpts[b][:, now], gt_areas[b], wh)
                host = torch.empty(cost.shape, dtype=cost.dtype, pin_memory=True)
                host.copy_(cost, non_blocking=True)
                per_clip.append(host)
            pending.append(per_clip)
        torch.cuda.current_stream(dev).synchronize()
        solved, flat = [], []
        for per_clip in pending:
            out = []
            for host in per_clip:
                rows, cols = ([], []) if host is None else linear_sum_assignment(host.numpy())
                out.append((len(flat), len(rows)))
                flat.extend(rows)
                flat.extend(cols)
            solved.append(out)
        # one host->device copy for all index lists
        idx = torch.tensor(flat, dtype=torch.long).pin_memory().to(dev, non_blocking=True) if flat else \
            torch.zeros(0, dtype=torch.long, device=dev)
        return [[(idx[o:o + n], idx[o + n:o + 2 * n]) for o, n in out] for out in solved]

    def cls_loss(self, cls, match):
        """Focal classification loss of one stage; cls (Bc, N, 1)."""
        labels = torch.zeros_like(cls)
        for b, (rows, _) in enumerate(match):
            labels[b, rows] = 1.0
        n_pos = float(sum(int(r.shape[0]) for r, _ in match))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            n_pos = reduce_mean(cls.new_tensor([n_pos])).clamp(min=1)
        else:
            n_pos = max(n_pos, 1.0)
        p = cls.sigmoid()
        focal = F.binary_cross_entropy_with_logits(cls, labels, reduction='none') * \
            (labels * 0.25 * (1 - p) ** 2 + (1 - labels) * 0.75 * p ** 2)
        return 0.5 * focal.sum() / n_pos

    def kpt_terms(self, kpt, sigma, match, gt_kpts):
        """(pred, sigma, target, weight) of one stage's matched poses, each (N, K, 2);
        kpt (Bc, T, N, 2K), sigma (Bc, N, 2K)."""
        K, now = self.K, self.T // 2
        preds, sigmas, tgts, wts = [], [], [], []
        for b, (rows, cols) in enumerate(match):
            preds.append(kpt[b, now, rows].view(-1, K, 2))
            sigmas.append(sigma[b, rows].view(-1, K, 2))
            tgts.append(gt_kpts[b][cols, now, :, :2])
            wts.append((gt_kpts[b][cols, now, :, 2:] > 0).float().expand(-1, -1, 2))
        return tuple(torch.cat(t) for t in (preds, sigmas, tgts, wts))

    def refine(self, memory, mask_flat, shapes, lsi, kpt_last, matches, gt_kpts):
        """Joint decoder over the matched persons: K keypoint queries per person.
        Returns the keypoint-loss terms of its layers for `rle_all`."""
        T, K = self.T, self.K
        poses, tgts, wts = [], [], []
        for b, (rows, cols) in enumerate(matches):
            poses.append(kpt_last[b][:, rows])                                          # (T, g, 2K)
            tgts.append(gt_kpts[b][cols, T // 2, :, :2])
            wts.append((gt_kpts[b][cols, T // 2, :, 2:] > 0).float().expand(-1, -1, 2))
        poses = torch.cat(poses, 1)
        group_sizes = [int(rows.shape[0]) for rows, _ in matches]
        G = sum(group_sizes)
        names = ['d%d.loss_kpt_refine' % lid for lid in range(len(self.refine_decoder))]
        if G == 0:
            zero = sum(p.sum() for p in self.refine_decoder.parameters()) * 0 + \
                sum(p.sum() for p in self.refine_kpt.parameters()) * 0 + \
                sum(p.sum() for p in self.refine_sigma.parameters()) * 0 + \
                self.refine_query_embedding.weight.sum() * 0
            empty = zero.new_zeros(1, K, 2)
            # one all-zero-weight row per layer keeps the parameters in the autograd graph (DDP)
            return [(n, empty + zero, empty + 1, empty, empty) for n in names]
        ref = poses.detach().reshape(T * G, K, 2).contiguous()                          # frame-major (T*G, K, 2)
        tgt, wt = torch.cat(tgts), torch.cat(wts)
        # the reference gathers memory[:, img_inds] -> (S, G, T, 256) here; the persons of a clip
        # share its tokens instead (value_group_sizes), so nothing is gathered or re-projected
        x = memory.transpose(0, 1)                                                      # (Bc*T, S, 256)
        graphed = self._graphed
        if graphed is not None and self.training and torch.is_grad_enabled():
            outs = graphed['joint'](x, mask_flat, shapes, lsi, ref, static=tuple(group_sizes))
        else:
            outs = self._stage_joint(x, mask_flat, shapes, lsi, ref, static=tuple(group_sizes))
        return [(n, outs[2 * lid][G:2 * G], outs[2 * lid + 1], tgt, wt) for lid, n in enumerate(names)]

    def _stage_joint(self, x, mask_flat, shapes, lsi, ref, static):
        """The joint decoder layers for G matched persons (static = persons per clip):
        x (Bc*T, S, 256), ref (T*G, K, 2) -> (refined keypoints (T*G, K, 2), sigma (G, K, 2)) per layer.
        Shapes follow G, so a graphed run keeps one graph per distinct `static`."""
        T, K = self.T, self.K
        S, G = x.shape[1], ref.shape[0] // T
        q_pos, q = self.refine_query_embedding.weight.split(256, 1)
        query = q[None].expand(G, -1, -1).permute(1, 0, 2)                              # (K, G, 256)
        q_pos = q_pos[None].expand(G, -1, -1).permute(1, 0, 2)
        mem = x.reshape(-1, T, S, 256).permute(2, 0, 1, 3)                              # (S, Bc, T, 256) view
        mask = mask_flat.view(-1, T, S)                                                 # (Bc, T, S)
        outs = []
        for lid, layer in enumerate(self.refine_decoder):
            ref_in = ref[:, :, None, :].expand(-1, -1, 4, -1)                           # (T*G, K, L, 2)
            query = layer(query, q_pos, value=mem, key_padding_mask=mask, reference_points=ref_in,
                          spatial_shapes=shapes, level_start_index=lsi, value_group_sizes=list(static))
            o = query.permute(1, 0, 2)                                                  # (G, K, 256)
            deltas = torch.cat([self.refine_kpt[t][lid](o) for t in range(T)], 0)       # (T*G, K, 2)
            ref = (deltas + inverse_sigmoid(ref)).sigmoid()
            outs += [ref, self.refine_sigma[lid](o).sigmoid()]
        return tuple(outs)


def build_optimizer(model):
    """AdamW 2e-5 / wd 1e-4 with 0.1x lr on backbone and sampling_offsets
    (config lines 140-150)."""
    slow, fast = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        backbone = name.startswith(('stem', 'layer1', 'layer2', 'layer3', 'layer4'))
        (slow if backbone or 'sampling_offsets' in name else fast).append(p)
    on_gpu = all(p.is_cuda for p in fast + slow)
    return torch.optim.AdamW([{'params': fast, 'lr': 2e-5}, {'params': slow, 'lr': 2e-6}],
                             weight_decay=1e-4, fused=on_gpu)   # one multi-tensor kernel per group


def synthetic_clip_batch(clips, device, seed, height=800, width=1333, num_frames=3, num_keypoints=15):
    """PoseTrack-shaped synthetic input: images ~N(0,1); per clip 1-10 persons with
    K keypoints uniform in a random box, small drift between frames."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    images = torch.randn(clips, num_frames, 3, height, width, generator=g)
    kpts, areas = [], []
    for _ in range(clips):
        n = int(torch.randint(1, 11, (1,), generator=g))
        centre = torch.rand(n, 1, 1, 2, generator=g) * 0.6 + 0.2
        size = torch.rand(n, 1, 1, 2, generator=g) * 0.3 + 0.1
        xy = centre + (torch.rand(n, 1, num_keypoints, 2, generator=g) - 0.5) * size
        xy = (xy + torch.randn(n, num_frames, 1, 2, generator=g) * 0.01).clamp(0.01, 0.99)
        vis = (torch.rand(n, num_frames, num_keypoints, 1, generator=g) > 0.15).float() * 2
        kpts.append(torch.cat([xy, vis], -1).to(device))
        areas.append((size[:, 0, 0, 0] * size[:, 0, 0, 1]).to(device))
    return images.to(device), kpts, areas


class FlatGradients(object):
    """Every trainable parameter's .grad is a view into ONE flat fp32 buffer, laid out in the
    order in which the backward pass completes the gradients: [joint decoder + pose decoder +
    heads | encoder | backbone top | backbone trunk].  The clip-sharded step exchanges the buffer bucket by bucket
    with NCCL all-reduces on a side stream, each launched from an autograd hook the moment the
    stage before it (in backward order) has finished, so the exchange of the decoders', the
    encoder's and the backbone top's gradients runs underneath the rest of the backward and only
    the last bucket (layer2-3, 17 % of the bytes) is exposed.  No hooks or buckets on the parameters themselves, which keeps the step
    compatible with the CUDA-graphed stages (autograd accumulates into the views in place).
    The reference's equivalent is MMDistributedDataParallel, i.e. torch DDP's bucketed,
    backward-overlapped all-reduce (opera/apis/train.py:153-162,
    third_party/mmcv/mmcv/parallel/distributed.py:11-85)."""

    def __init__(self, model, overlap=True):
        buckets = model.gradient_buckets() if hasattr(model, 'gradient_buckets') else \
            [[p for p in model.parameters() if p.requires_grad]]
        self.params = [p for b in buckets for p in b]
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=self.params[0].dtype, device=self.params[0].device)
        offset, self.ranges = 0, []
        for b in buckets:
            start = offset
            for p in b:
                p.grad = self.flat[offset:offset + p.numel()].view_as(p)
                offset += p.numel()
            self.ranges.append((start, offset))
        self.overlap = overlap
        self._side = None
        self._pending, self._launched = [], 0
        self.exposed_events = None          # (start, end) CUDA events around the exposed wait of the last step
        if overlap and hasattr(model, 'grad_exchange'):
            model.grad_exchange = self

    def _distributed(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def zero(self):
        self.flat.zero_()
        self._pending, self._launched = [], 0

    def launch_through(self, n_buckets):
        """All-reduce (mean) buckets [launched, n_buckets) on the side stream, behind everything
        the current stream has been given so far.  Called from the autograd hooks."""
        if not (self.overlap and self._distributed()):
            return
        on_gpu = self.flat.is_cuda
        if on_gpu:
            cur = torch.cuda.current_stream(self.flat.device)
            if self._side is None:
                self._side = torch.cuda.Stream(self.flat.device)
        while self._launched < min(n_buckets, len(self.ranges)):
            a, b = self.ranges[self._launched]
            self._launched += 1
            if b == a:
                continue
            if on_gpu:
                ready = torch.cuda.Event()
                ready.record(cur)
                self._side.wait_event(ready)
                with torch.cuda.stream(self._side):
                    self._pending.append(dist.all_reduce(self.flat[a:b], op=dist.ReduceOp.AVG,
                                                         async_op=True))
            else:       # CPU tensors (gloo, tests): no streams, no AVG — sum now, divide at the end
                self._pending.append(dist.all_reduce(self.flat[a:b], async_op=True))

    def all_reduce_mean(self):
        """Finish the exchange: whatever has not been launched yet goes now (exposed), then the
        current stream waits for the side stream."""
        if not self._distributed():
            return
        dev = self.flat.device
        if not self.flat.is_cuda:
            if self.overlap:
                self.launch_through(len(self.ranges))
                for work in self._pending:
                    work.wait()
                self._pending = []
            else:
                dist.all_reduce(self.flat)
            self.flat.div_(dist.get_world_size())
            return
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.current_stream(dev))
        if self.overlap:
            self.launch_through(len(self.ranges))
            for work in self._pending:
                work.wait()                      # current stream waits for the collective
            self._pending = []
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        ev1.record(torch.cuda.current_stream(dev))
        self.exposed_events = (ev0, ev1)


def train_step(model, optimizer, images, gt_kpts, gt_areas, ddp_model=None, flat_grads=None):
    """forward + backward + gradient all-reduce + grad-clip 0.1 + AdamW step.
    Gradient exchange: `flat_grads` (FlatGradients: one all-reduce of a flat bucket) or
    `ddp_model` (torch DDP hooks; not combinable with enable_graphs)."""
    net = ddp_model if ddp_model is not None else model
    if getattr(model, '_graphed', None) is not None:
        from . import graphs
        graphs.refresh_seed(images.device)        # new epilogue-dropout masks for this step's replays
    losses = net(images, gt_kpts, gt_areas)
    loss = sum(losses.values())
    if flat_grads is not None:
        flat_grads.zero()
    else:
        optimizer.zero_grad(set_to_none=True)
    if getattr(model, '_graphed', None) is not None:
        from . import graphs
        with graphs.quiet_accumulate_grad_stream_warning():
            loss.backward()
    else:
        loss.backward()
    if flat_grads is not None:
        flat_grads.all_reduce_mean()
        torch.nn.utils.clip_grad_norm_(flat_grads.params, 0.1)
    else:
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 0.1)
    optimizer.step()
    return loss.detach()


PaveNetR50.forward = PaveNetR50.forward_train
