"""CAV-MAE audio encoder and Modal3Classifier (reference models/cav_mae.py:69-151,337-351 and models/basic_model.py:202-275)
for the three-modality path --lorb m3ae --modal3 --gs_flag (BASELINE.json configs[3], IEMOCAP-shaped).

`CAVMAEFT.forward_feat(a, None, 'a')` is the only entry the path uses: spectrogram [B, T, 128] -> 16x16 patches (a stride-16
convolution = a Linear over 256-value patches) + learned position / modality embeddings -> 11 audio blocks + 1 shared block
read through its audio LayerNorms -> norm_a. The blocks are pre-LN ViT blocks whose attention / MLP classes the reference
imports from the un-vendored `timm==0.4.5` (cav_mae.py:15-16; `Attention`: qkv Linear with bias, softmax(q k^T / sqrt(d)) v,
proj Linear; `Mlp`: fc1 -> exact GELU -> fc2): the same arithmetic as the m3ae block without a key mask, so they run on the
same fused node (`m3ae._BlockFn`): tcgen05 GEMMs, fused attention, fused LayerNorm / GELU kernels.

State-dict keys and parameter creation order follow the reference (including the visual branch the audio mode never
touches: `patch_embed_v`, `blocks_v`, `norm_v`, ...), so its checkpoints load and seeding reproduces its initial weights.
Parameters the audio mode does not use never receive gradients (`hot_parameters()` lists the ones that do), exactly as in
the reference where the optimiser skips `grad is None`.
"""
import numpy as np
import torch
import torch.nn as nn

from . import m3ae
from .fusion_modules import ConcatFusion3
from .m3ae import MaskedMultimodalAutoencoder, NativeLinear, _BlockFn, _LayerNormFn


def sincos_2d_rect(embed_dim, grid_h, grid_w):
    """[grid_h * grid_w, embed_dim] (cav_mae.py:19-66): float64 sin-cos tables; first half of the channels from the column
    index, second half from the row (the reference's meshgrid puts w first)."""
    def one(dim, pos):
        omega = np.arange(dim // 2, dtype=float)
        omega /= dim / 2.
        omega = 1. / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)
    gw, gh = np.meshgrid(np.arange(grid_w, dtype=np.float32), np.arange(grid_h, dtype=np.float32))
    return np.concatenate([one(embed_dim // 2, gw), one(embed_dim // 2, gh)], axis=1)


class PatchEmbed(nn.Module):                           # cav_mae.py:69-84
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        """x [B, C, H, W] -> [B, (H/p)(W/p), embed_dim]: the stride-p convolution as a Linear over (c, p1, p2) patches."""
        B, C, H, W = x.shape
        p = self.patch_size[0]
        patches = x.reshape(B, C, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), C * p * p)
        w = self.proj.weight.view(self.proj.weight.shape[0], -1)
        return m3ae._LinearFn.apply(patches.reshape(-1, C * p * p).float().contiguous(), w, self.proj.bias).view(
            B, -1, w.shape[0])


class _Attention(nn.Module):                           # timm 0.4.5 vision_transformer.Attention (parameters only)
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = NativeLinear(dim, dim * 3, bias=qkv_bias)
        self.proj = NativeLinear(dim, dim)


class _Mlp(nn.Module):                                 # timm 0.4.5 vision_transformer.Mlp (parameters only)
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = NativeLinear(dim, hidden)
        self.fc2 = NativeLinear(hidden, dim)


class Block(nn.Module):                                # cav_mae.py:86-113
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=True):
        super().__init__()
        self.norm1, self.norm1_a, self.norm1_v = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2, self.norm2_a, self.norm2_v = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def norms(self, modality):
        return {None: (self.norm1, self.norm2), "a": (self.norm1_a, self.norm2_a), "v": (self.norm1_v, self.norm2_v)}[modality]

    def forward(self, x, modality=None):
        if not x.is_cuda:
            raise RuntimeError("mla_b200 CAV-MAE encoder runs on CUDA only (no CPU fallback); got %s" % x.device)
        n1, n2 = self.norms(modality)
        a, m = self.attn, self.mlp
        return _BlockFn.apply(x, None, a.num_heads, a.scale, n1.eps, n2.eps, n1.weight, n1.bias, a.qkv.weight, a.qkv.bias,
                              a.proj.weight, a.proj.bias, n2.weight, n2.bias, m.fc1.weight, m.fc1.bias, m.fc2.weight,
                              m.fc2.bias)


class CAVMAEFT(nn.Module):                             # cav_mae.py:116-185 (constructor), :337-351 (audio features)
    def __init__(self, label_dim, img_size=224, audio_length=1024, patch_size=16, in_chans=3, embed_dim=768,
                 modality_specific_depth=11, num_heads=12, mlp_ratio=4., tr_pos=True):
        super().__init__()
        self.patch_embed_a = PatchEmbed(img_size, patch_size, 1, embed_dim)
        self.patch_embed_v = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.patch_embed_a.num_patches = int(audio_length * 128 / 256)
        self.modality_a = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.modality_v = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed_a = nn.Parameter(torch.zeros(1, self.patch_embed_a.num_patches, embed_dim), requires_grad=tr_pos)
        self.pos_embed_v = nn.Parameter(torch.zeros(1, self.patch_embed_v.num_patches, embed_dim), requires_grad=tr_pos)
        mk = lambda: Block(embed_dim, num_heads, mlp_ratio, qkv_bias=True)      # noqa: E731
        self.blocks_a = nn.ModuleList([mk() for _ in range(modality_specific_depth)])
        self.blocks_v = nn.ModuleList([mk() for _ in range(modality_specific_depth)])
        self.blocks_u = nn.ModuleList([mk() for _ in range(12 - modality_specific_depth)])
        self.norm_a = nn.LayerNorm(embed_dim)
        self.norm_v = nn.LayerNorm(embed_dim)
        self.initialize_weights()

    def initialize_weights(self):                      # cav_mae.py:160-185
        D = self.pos_embed_a.shape[-1]
        na, nv = self.patch_embed_a.num_patches, self.patch_embed_v.num_patches
        self.pos_embed_a.data.copy_(torch.from_numpy(sincos_2d_rect(D, 8, int(na / 8))).float().unsqueeze(0))
        self.pos_embed_v.data.copy_(torch.from_numpy(sincos_2d_rect(D, int(nv ** .5), int(nv ** .5))).float().unsqueeze(0))
        for pe in (self.patch_embed_a, self.patch_embed_v):
            w = pe.proj.weight.data
            nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.normal_(self.modality_a, std=.02)
        nn.init.normal_(self.modality_v, std=.02)

        def init(mod):
            if isinstance(mod, nn.Linear):
                nn.init.xavier_uniform_(mod.weight)
                if mod.bias is not None:
                    nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.LayerNorm):
                nn.init.constant_(mod.bias, 0)
                nn.init.constant_(mod.weight, 1.0)
        self.apply(init)

    def hot_parameters(self):
        """The parameters forward_feat(a, None, 'a') reads, i.e. the ones that receive gradients on this path."""
        ps = list(self.patch_embed_a.parameters()) + [self.modality_a]
        if self.pos_embed_a.requires_grad:
            ps.append(self.pos_embed_a)
        for blk in self.blocks_a:
            ps += [blk.norm1.weight, blk.norm1.bias] + list(blk.attn.parameters()) + [blk.norm2.weight, blk.norm2.bias] \
                + list(blk.mlp.parameters())
        for blk in self.blocks_u:
            ps += [blk.norm1_a.weight, blk.norm1_a.bias] + list(blk.attn.parameters()) \
                + [blk.norm2_a.weight, blk.norm2_a.bias] + list(blk.mlp.parameters())
        return ps + list(self.norm_a.parameters())

    def forward_feat(self, a, v=None, mode="a"):
        if mode != "a":
            raise NotImplementedError("mla_b200 implements the audio mode of CAVMAEFT.forward_feat (the MLA modal3 path)")
        a = a.unsqueeze(1).transpose(2, 3)             # [B, T, 128] -> [B, 1, 128, T]
        a = self.patch_embed_a(a) + self.pos_embed_a + self.modality_a
        for blk in self.blocks_a:
            a = blk(a)
        for blk in self.blocks_u:
            a = blk(a, "a")
        return _LayerNormFn.apply(a, self.norm_a.weight, self.norm_a.bias, self.norm_a.eps)


class Modal3Classifier(nn.Module):
    """basic_model.py:202-275. forward(token [B,1,L], padding_mask [B,1,L], visual [B,3,H,W], audio [B,T,128]) ->
    (a, v, t), each [B, 768]."""

    def __init__(self, args, model_config=None, text_vocab_size=30522, audio_kwargs=None):
        super().__init__()
        if args.dataset != "IEMOCAP":
            raise NotImplementedError("Incorrect dataset name {}".format(args.dataset))
        if args.fusion_method != "concat":
            raise NotImplementedError("mla_b200 implements the concat head only")
        if getattr(args, "modulation", "Normal") == "QMF":
            raise NotImplementedError("QMF is outside the MLA hot path (SURVEY.md section 2)")
        n_classes = 4
        model_config = dict(model_config or {"model_type": "base"})
        audio_kwargs = dict(audio_kwargs or {})
        emb = m3ae._SIZES[model_config["model_type"]][0] if model_config.get("model_type") else model_config["emb_dim"]
        self.fusion_module = ConcatFusion3(input_dim=emb if args.gs_flag else 3 * emb, output_dim=n_classes)   # basic_model.py:218 / 221
        self.mae_a = CAVMAEFT(n_classes, **audio_kwargs)                                 # basic_model.py:231
        self.mae_v = MaskedMultimodalAutoencoder(text_vocab_size, model_config)
        self.mae_t = MaskedMultimodalAutoencoder(text_vocab_size, model_config)
        # basic_model.py:234-242 loads pretrained encoders from placeholder paths; optional here, non-strict like there
        for enc, key in ((self.mae_a, "cav_ckpt_audio"), (self.mae_v, "m3ae_ckpt"), (self.mae_t, "m3ae_ckpt")):
            path = getattr(args, key, None)
            if path:
                enc.load_state_dict(torch.load(path, map_location="cpu"), strict=False)
        self.args = args

    def forward(self, token, padding_mask, visual, audio):
        B, Cc, H, W = visual.shape
        p = 16
        patches = visual.reshape(B, Cc, H // p, p, W // p, p).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // p) * (W // p), Cc * p * p)
        a = self.mae_a.forward_feat(audio, None, "a")
        t = self.mae_t.forward_representation(None, token.squeeze(1), padding_mask.squeeze(1))
        v = self.mae_v.forward_representation(patches, None, None)
        return a.mean(dim=1), v.mean(dim=1), t.mean(dim=1)
