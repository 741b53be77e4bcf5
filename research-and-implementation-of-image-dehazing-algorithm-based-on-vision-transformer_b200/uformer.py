"""Host-side model graph that CALLS the hot path: Uformer (My_model_1.py:955-1230).

Out of the hot-path scope: InputProj / OutputProj / Downsample /
Upsample are stock PyTorch convolutions (cuDNN) with the reference's parameters; only the 18
LeWinTransformerBlocks run on the sm_100a kernels.  The convolutions are fed channels_last VIEWS of the
token-major residual stream, so the reference's NCHW <-> token transposes disappear (same arithmetic).  This file exists because the reference source
cannot travel to the GPU box; it keeps the reference's constructor arguments, forward signature and
the 488-key state_dict layout (tests/golden/uformer32_state_dict_keys.txt) so checkpoints load
strictly (utils/model_utils.py:28-40).

One deliberate host-side difference: the 18 ``index_sample`` draws (attn.py:91) of a forward are made
up-front, in module order, from the same CPU generator — the RNG stream is identical to the
reference's, but the 18 small host->device copies collapse into one.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .modules import LeWinTransformerBlock, draw_index_sample


def _tokens_to_nchw(x, hw=None):
    """[B, L, C] token-major -> [B, C, H, W] VIEW with channels_last strides (no copy): the residual stream is already
    NHWC, so cuDNN's NHWC kernels consume it directly (the reference transposes to NCHW and back around every
    convolution, My_model_1.py:620-621, 646-647, 679, 719).  hw: explicit (H, W) of a non-square map (row bands)."""
    B, L, C = x.shape
    H, W = hw if hw is not None else (int(math.sqrt(L)),) * 2
    return x.reshape(B, H, W, C).permute(0, 3, 1, 2)


def _nchw_to_tokens(y):
    """[B, C, H, W] (channels_last after an NHWC convolution) -> [B, H*W, C]; a copy only if y is not channels_last."""
    B, C, H, W = y.shape
    return y.permute(0, 2, 3, 1).reshape(B, H * W, C)


def _autocast_bf16():
    """bf16 CUDA autocast active (torch >= 2.4 API, older fallback)."""
    if not torch.is_autocast_enabled():
        return False
    try:
        return torch.get_autocast_dtype("cuda") == torch.bfloat16
    except (AttributeError, TypeError):
        return torch.get_autocast_gpu_dtype() == torch.bfloat16


def _cl(w):
    return w.contiguous(memory_format=torch.channels_last)


class Downsample(nn.Module):
    """My_model_1.py:606-622: 4x4 stride-2 conv on the token map."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channel, out_channel, kernel_size=4, stride=2, padding=1))
        self.in_channel, self.out_channel = in_channel, out_channel

    def forward(self, x, hw=None, pad_h=True):
        """hw: explicit (H, W) of a non-square token map; pad_h=False: the rows already carry their halo (canvas row bands)."""
        c = self.conv[0]
        B, L, C = x.shape
        H, W = hw if hw is not None else (int(math.sqrt(L)),) * 2
        if not (torch.is_grad_enabled() and (x.requires_grad or c.weight.requires_grad)):
            from . import ops
            xb = x if x.dtype == torch.bfloat16 else (x.to(torch.bfloat16) if x.is_cuda and _autocast_bf16() else None)
            if xb is not None and H % 2 == 0 and ops.downsample_supported(xb, C, W):
                # bf16 inference: implicit GEMM on tcgen05 with the bias fused (lewin_downsample_fwd_bf16)
                return ops.lewin_downsample(xb, c.weight, c.bias, B=B, H=H, W=W, pad_h=pad_h)
        y = torch.nn.functional.conv2d(_tokens_to_nchw(x, (H, W)), _cl(c.weight), c.bias, stride=2, padding=(1 if pad_h else 0, 1))
        return _nchw_to_tokens(y)


class Upsample(nn.Module):
    """My_model_1.py:633-648: 2x2 stride-2 transposed conv."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.deconv = nn.Sequential(nn.ConvTranspose2d(in_channel, out_channel, kernel_size=2, stride=2))
        self.in_channel, self.out_channel = in_channel, out_channel

    def _gemm_path(self, x):
        """bf16 inference: the non-overlapping transposed convolution is a token GEMM with a pixel-shuffle row address
        (lewin_upsample_fwd_bf16); training / fp32 keep cuDNN."""
        from . import ops
        d = self.deconv[0]
        if torch.is_grad_enabled() and (x.requires_grad or d.weight.requires_grad):
            return False
        xb = x if x.dtype == torch.bfloat16 else None
        if xb is None and x.is_cuda and _autocast_bf16():
            xb = x.to(torch.bfloat16)
        if xb is None or not ops.upsample_supported(xb, self.in_channel, self.out_channel, x.shape[0] * x.shape[1]):
            return False
        return xb

    def forward(self, x, skip=None, hw=None, cat_buf=None):
        """``skip`` given: returns torch.cat([up(x), skip], -1) (My_model_1.py:1189-1204) with the up half written in place.
        hw: explicit (H, W) of a non-square token map (canvas row bands); default: square, as the reference assumes.
        cat_buf: a [B, 4L, Cout + Cskip] buffer whose right half IS ``skip`` (the encoder wrote its output there): no copy."""
        d = self.deconv[0]
        xb = self._gemm_path(x)
        if xb is not False:
            from . import ops
            B, L, _ = x.shape
            H, W = hw if hw is not None else (int(math.sqrt(L)),) * 2
            if skip is None:
                return ops.lewin_upsample(xb, d.weight, d.bias, B=B, H=H, W=W)
            C = self.out_channel
            in_place = (cat_buf is not None and cat_buf.dtype == xb.dtype and tuple(cat_buf.shape) == (B, 4 * L, C + skip.shape[-1]) and
                        cat_buf.is_contiguous() and skip.data_ptr() == cat_buf[..., C:].data_ptr() and skip.stride(-2) == cat_buf.stride(-2))
            buf = cat_buf if in_place else torch.empty((B, 4 * L, C + skip.shape[-1]), dtype=xb.dtype, device=x.device)
            ops.lewin_upsample(xb, d.weight, d.bias, B=B, H=H, W=W, out=buf)
            if not in_place:
                buf[..., C:] = skip
            return buf
        y = torch.nn.functional.conv_transpose2d(_tokens_to_nchw(x, hw), _cl(d.weight), d.bias, stride=2)
        y = _nchw_to_tokens(y)
        return y if skip is None else torch.cat([y, skip], -1)


class InputProj(nn.Module):
    """My_model_1.py:659-682: 3x3 conv + LeakyReLU -> tokens."""

    def __init__(self, in_channel=3, out_channel=64, kernel_size=3, stride=1, norm_layer=None, act_layer=nn.LeakyReLU):
        super().__init__()
        self.proj = nn.Sequential(
            nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=stride, padding=kernel_size // 2),
            act_layer(inplace=True))
        self.norm = norm_layer(out_channel) if norm_layer is not None else None
        self.in_channel, self.out_channel = in_channel, out_channel

    def forward(self, x):
        conv, act = self.proj[0], self.proj[1]
        fast = (x.is_cuda and x.dtype == torch.float32 and _autocast_bf16() and isinstance(act, nn.LeakyReLU) and
                conv.in_channels <= 4 and conv.out_channels in (32, 64) and conv.stride == (1, 1) and
                not (torch.is_grad_enabled() and (x.requires_grad or conv.weight.requires_grad)))
        if fast:        # bf16 inference: convolution + bias + LeakyReLU in one kernel (lewin_input_proj_fwd_bf16)
            from . import ops
            x = ops.lewin_input_proj(x, conv.weight, conv.bias, act.negative_slope)
        else:
            x = _nchw_to_tokens(self.proj(x.contiguous(memory_format=torch.channels_last)))
        return self.norm(x) if self.norm is not None else x


class OutputProj(nn.Module):
    """My_model_1.py:696-723: tokens -> 3x3 conv."""

    def __init__(self, in_channel=64, out_channel=3, kernel_size=3, stride=1, norm_layer=None, act_layer=None):
        super().__init__()
        self.proj = nn.Sequential(
            nn.Conv2d(in_channel, out_channel, kernel_size=3, stride=stride, padding=kernel_size // 2))
        self.norm = norm_layer(out_channel) if norm_layer is not None else None
        self.in_channel, self.out_channel = in_channel, out_channel

    def forward(self, x, residual=None, hw=None, pad_h=True):
        """residual: optional fp32 image added to the result (the `x + y` of Uformer.forward, fused into the kernel's store);
        hw / pad_h as Downsample.forward."""
        c = self.proj[0]
        B, L, C = x.shape
        H, W = hw if hw is not None else (int(math.sqrt(L)),) * 2
        if self.norm is None and (residual is None or residual.dtype == torch.float32) and not (
                torch.is_grad_enabled() and (x.requires_grad or c.weight.requires_grad)):
            from . import ops
            xb = x if x.dtype == torch.bfloat16 else (x.to(torch.bfloat16) if x.is_cuda and _autocast_bf16() else None)
            if xb is not None and c.stride == (1, 1) and ops.output_proj_supported(xb, C, c.out_channels, W):
                # bf16 inference: implicit GEMM on tcgen05, bias + residual image fused (lewin_output_proj_fwd_bf16)
                return ops.lewin_output_proj(xb, c.weight, c.bias, B=B, H=H, W=W, residual=residual, pad_h=pad_h)
        y = torch.nn.functional.conv2d(_tokens_to_nchw(x, (H, W)), c.weight, c.bias, stride=c.stride, padding=(1 if pad_h else 0, 1))
        y = self.norm(y) if self.norm is not None else y
        return y if residual is None else residual + y.to(residual.dtype)


class BasicUformerLayer(nn.Module):
    """My_model_1.py:894-946: ``depth`` LeWin blocks, shift 0 / win//2 alternating."""

    def __init__(self, dim, output_dim, input_resolution, depth, num_heads, win_size, mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, use_checkpoint=False,
                 token_projection='linear', token_mlp='ffn', se_layer=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth = dim, input_resolution, depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            LeWinTransformerBlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, win_size=win_size,
                                  shift_size=0 if (i % 2 == 0) else win_size // 2, mlp_ratio=mlp_ratio,
                                  qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                  drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                  norm_layer=norm_layer, token_projection=token_projection, token_mlp=token_mlp,
                                  se_layer=se_layer)
            for i in range(depth)])

    def forward(self, x, mask=None, index_samples=None, out=None):
        """out: optional destination view for the layer's result (handed to the last block)."""
        last = len(self.blocks) - 1
        for i, blk in enumerate(self.blocks):
            x = blk(x, mask, None if index_samples is None else index_samples[i], **({"out": out} if (out is not None and i == last) else {}))
        return x


class Uformer(nn.Module):
    """My_model_1.py:955-1207.  ``forward(x[B,3,H,W], mask=None) -> [B,3,H,W]`` (H == W, multiple of 128)."""

    def __init__(self, img_size=128, in_chans=3, embed_dim=32, depths=[2, 2, 2, 2, 2, 2, 2, 2, 2],
                 num_heads=[1, 2, 4, 8, 16, 16, 8, 4, 2], win_size=8, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0.1, norm_layer=nn.LayerNorm, patch_norm=True,
                 use_checkpoint=False, token_projection='linear', token_mlp='ffn', se_layer=False,
                 dowsample=Downsample, upsample=Upsample, **kwargs):
        super().__init__()
        if token_mlp != 'leff' or token_projection != 'linear':
            raise NotImplementedError(
                "lewin_b200.Uformer implements the configuration the reference instantiates (utils/model_utils.py:94: "
                f"token_projection='linear', token_mlp='leff'); got token_projection={token_projection!r}, token_mlp={token_mlp!r} "
                "(the constructor keeps the reference's default token_mlp='ffn' in its signature, so pass token_mlp='leff')")
        if embed_dim not in (32, 64):
            raise NotImplementedError(
                f"lewin_b200.Uformer: embed_dim {embed_dim} is not built.  head_dim = embed_dim (My_model_1.py:962) must be 32, 64 or "
                "128 and the widest level (16 x embed_dim channels at the bottleneck) at most 1024 channels, so whole models exist "
                "for embed_dim 32 and 64; head_dim 128 is available at block level (LeWinTransformerBlock with dim <= 1024)")
        self.num_enc_layers = len(depths) // 2
        self.num_dec_layers = len(depths) // 2
        self.embed_dim, self.patch_norm, self.mlp_ratio = embed_dim, patch_norm, mlp_ratio
        self.token_projection, self.mlp, self.win_size, self.reso = token_projection, token_mlp, win_size, img_size
        self.depths = list(depths)
        self.pos_drop = nn.Dropout(p=drop_rate)

        enc_dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths[:self.num_enc_layers]))]
        conv_dpr = [drop_path_rate] * depths[4]
        dec_dpr = enc_dpr[::-1]

        self.input_proj = InputProj(in_channel=in_chans, out_channel=embed_dim, kernel_size=3, stride=1,
                                    act_layer=nn.LeakyReLU)
        self.output_proj = OutputProj(in_channel=2 * embed_dim, out_channel=in_chans, kernel_size=3, stride=1)

        def layer(mult, res_div, i, dpr):
            return BasicUformerLayer(dim=embed_dim * mult, output_dim=embed_dim * mult,
                                     input_resolution=(img_size // res_div, img_size // res_div), depth=depths[i],
                                     num_heads=num_heads[i], win_size=win_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr,
                                     norm_layer=norm_layer, use_checkpoint=use_checkpoint,
                                     token_projection=token_projection, token_mlp=token_mlp, se_layer=se_layer)

        d = depths
        self.encoderlayer_0 = layer(1, 1, 0, enc_dpr[sum(d[:0]):sum(d[:1])])
        self.dowsample_0 = dowsample(embed_dim, embed_dim * 2)
        self.encoderlayer_1 = layer(2, 2, 1, enc_dpr[sum(d[:1]):sum(d[:2])])
        self.dowsample_1 = dowsample(embed_dim * 2, embed_dim * 4)
        self.encoderlayer_2 = layer(4, 4, 2, enc_dpr[sum(d[:2]):sum(d[:3])])
        self.dowsample_2 = dowsample(embed_dim * 4, embed_dim * 8)
        self.encoderlayer_3 = layer(8, 8, 3, enc_dpr[sum(d[:3]):sum(d[:4])])
        self.dowsample_3 = dowsample(embed_dim * 8, embed_dim * 16)
        self.conv = layer(16, 16, 4, conv_dpr)
        self.upsample_0 = upsample(embed_dim * 16, embed_dim * 8)
        self.decoderlayer_0 = layer(16, 8, 5, dec_dpr[:d[5]])
        self.upsample_1 = upsample(embed_dim * 16, embed_dim * 4)
        self.decoderlayer_1 = layer(8, 4, 6, dec_dpr[sum(d[5:6]):sum(d[5:7])])
        self.upsample_2 = upsample(embed_dim * 8, embed_dim * 2)
        self.decoderlayer_2 = layer(4, 2, 7, dec_dpr[sum(d[5:7]):sum(d[5:8])])
        self.upsample_3 = upsample(embed_dim * 4, embed_dim)
        self.decoderlayer_3 = layer(2, 1, 8, dec_dpr[sum(d[5:8]):sum(d[5:9])])
        self.apply(self._init_weights)

    def _init_weights(self, m):
        """My_model_1.py:1149-1156."""
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def extra_repr(self) -> str:
        return (f"embed_dim={self.embed_dim}, token_projection={self.token_projection}, "
                f"token_mlp={self.mlp},win_size={self.win_size}")

    def draw_index_samples(self):
        """The 18 (sum(depths)) draws of attn.py:91 in module execution order; one stacked CPU tensor."""
        return torch.stack([draw_index_sample(64, 64) for _ in range(sum(self.depths))])

    def forward(self, x, mask=None, index_samples=None):
        if index_samples is None:
            index_samples = self.draw_index_samples()
        idx = index_samples.to(device=x.device, dtype=torch.int32, non_blocking=True)
        d = self.depths
        offs = [sum(d[:i]) for i in range(len(d) + 1)]
        sl = lambda i: idx[offs[i]:offs[i + 1]]

        y = self.pos_drop(self.input_proj(x))
        # bf16 inference: every encoder level writes its result straight into the right half of the buffer the decoder's
        # torch.cat([up, skip], -1) (My_model_1.py:1189-1204) would build, so the four skip copies disappear
        cats = [None] * 4
        if y.is_cuda and y.dtype == torch.bfloat16 and not torch.is_grad_enabled():
            B, L0, E = y.shape
            cats = [torch.empty((B, L0 >> (2 * i), 2 * (E << i)), dtype=y.dtype, device=y.device) for i in range(4)]
        dst = lambda i: {} if cats[i] is None else {"out": cats[i][..., cats[i].shape[-1] // 2:]}
        conv0 = self.encoderlayer_0(y, mask, sl(0), **dst(0))
        pool0 = self.dowsample_0(conv0)
        conv1 = self.encoderlayer_1(pool0, mask, sl(1), **dst(1))
        pool1 = self.dowsample_1(conv1)
        conv2 = self.encoderlayer_2(pool1, mask, sl(2), **dst(2))
        pool2 = self.dowsample_2(conv2)
        conv3 = self.encoderlayer_3(pool2, mask, sl(3), **dst(3))
        pool3 = self.dowsample_3(conv3)
        conv4 = self.conv(pool3, mask, sl(4))
        deconv0 = self.decoderlayer_0(self.upsample_0(conv4, conv3, cat_buf=cats[3]), mask, sl(5))      # cat([up, skip], -1)
        deconv1 = self.decoderlayer_1(self.upsample_1(deconv0, conv2, cat_buf=cats[2]), mask, sl(6))
        deconv2 = self.decoderlayer_2(self.upsample_2(deconv1, conv1, cat_buf=cats[1]), mask, sl(7))
        deconv3 = self.decoderlayer_3(self.upsample_3(deconv2, conv0, cat_buf=cats[0]), mask, sl(8))
        return self.output_proj(deconv3, residual=x)                                   # x + y (My_model_1.py:1207)
