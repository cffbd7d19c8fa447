"""ORACLE — test infrastructure only (imported by tests/, never by the product).

CPU restatement of the HunyuanVideo path of the reference: `magcache_forward` of MagCache4HunyuanVideo/magcache_sample_video.py:29-160
(the patched `HYVideoDiffusionTransformer.forward`) and the transformer it drives.

PARITY UNPINNED for the block arithmetic: `MMDoubleStreamBlock`, `MMSingleStreamBlock`, `SingleTokenRefiner`, `TimestepEmbedder`,
`MLPEmbedder`, `FinalLayer`, `PatchEmbed`, `RMSNorm`, `modulate`, `apply_gate`, `apply_rotary_emb`, `get_cu_seqlens` live in `hyvideo`
(github Tencent/HunyuanVideo), which is not under /root/reference and is unpinned by it; what follows restates them from
SURVEY Appendix B.3 under hyvideo's attribute names. The controller / cache / counter statements ARE the reference's (cited inline)
and are pinned by tests/golden/masks.json.

dtype rule: the sampler runs the transformer in bf16 without autocast — every tensor is bf16 (`DT`); `exact()` switches to float64.
Variable-length attention: `get_cu_seqlens(text_mask, img_len)` makes two segments per sample, [image + valid text] and [padded text];
restated here as a block-diagonal mask. The padded text tokens therefore never influence an image token.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

DT = torch.bfloat16


class exact:
    def __enter__(self):
        global DT
        self._old, DT = DT, torch.float64

    def __exit__(self, *a):
        global DT
        DT = self._old


def _wide():
    return torch.float64 if DT == torch.float64 else torch.float32


def timestep_embedding(t, dim=256, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden, out=None):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(256, hidden), nn.SiLU(), nn.Linear(hidden, out or hidden))

    def forward(self, t):
        return self.mlp(timestep_embedding(t).type(self.mlp[0].weight.dtype))


class MLPEmbedder(nn.Module):
    def __init__(self, in_dim, hidden):
        super().__init__()
        self.in_layer, self.silu, self.out_layer = nn.Linear(in_dim, hidden), nn.SiLU(), nn.Linear(hidden, hidden)

    def forward(self, x):
        return self.out_layer(self.silu(self.in_layer(x)))


class TextProjection(nn.Module):
    def __init__(self, in_dim, hidden):
        super().__init__()
        self.linear_1, self.act_1, self.linear_2 = nn.Linear(in_dim, hidden), nn.SiLU(), nn.Linear(hidden, hidden)

    def forward(self, x):
        return self.linear_2(self.act_1(self.linear_1(x)))


class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps, self.weight = eps, nn.Parameter(torch.ones(dim))

    def forward(self, x):
        xf = x.to(_wide())
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x) * self.weight


class ModulateDiT(nn.Module):
    def __init__(self, hidden, factor):
        super().__init__()
        self.act, self.linear = nn.SiLU(), nn.Linear(hidden, factor * hidden)

    def forward(self, x):
        return self.linear(self.act(x))


class MLP(nn.Module):
    def __init__(self, dim, hidden, act):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(dim, hidden), nn.Linear(hidden, dim), act

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def apply_gate(x, gate):
    return x * gate.unsqueeze(1)


def _ln(x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)


def rotate_half(x):
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    return torch.stack([-x_imag, x_real], dim=-1).flatten(3)


def apply_rotary_emb(xq, xk, freqs_cis):
    """[B, L, H, D] layout (head_first=False); cos / sin [L, D] with every frequency repeated for its (real, imag) pair."""
    cos, sin = (t.to(_wide())[None, :, None, :] for t in freqs_cis)
    out = []
    for x in (xq, xk):
        xw = x.to(_wide())
        out.append((xw * cos + rotate_half(xw) * sin).type_as(x))
    return out


def segment_attention(q, k, v, seg_ids):
    """flash_attn_varlen_func over the segments of `get_cu_seqlens`: tokens attend inside their own segment only. [B, L, H, D]."""
    mask = (seg_ids[:, :, None] == seg_ids[:, None, :])[:, None]
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask)
    return o.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)


class MMDoubleStreamBlock(nn.Module):
    def __init__(self, hidden, heads, mlp_ratio=4):
        super().__init__()
        self.heads, hd = heads, hidden // heads
        act = nn.GELU(approximate="tanh")
        for p in ("img", "txt"):
            setattr(self, f"{p}_mod", ModulateDiT(hidden, 6))
            setattr(self, f"{p}_attn_qkv", nn.Linear(hidden, 3 * hidden))
            setattr(self, f"{p}_attn_q_norm", RMSNorm(hd))
            setattr(self, f"{p}_attn_k_norm", RMSNorm(hd))
            setattr(self, f"{p}_attn_proj", nn.Linear(hidden, hidden))
            setattr(self, f"{p}_mlp", MLP(hidden, mlp_ratio * hidden, act))

    def _qkv(self, p, x, shift, scale):
        qkv = getattr(self, f"{p}_attn_qkv")(modulate(_ln(x), shift, scale))
        b, l = x.shape[:2]
        q, k, v = qkv.view(b, l, 3, self.heads, -1).unbind(2)
        return getattr(self, f"{p}_attn_q_norm")(q).to(v), getattr(self, f"{p}_attn_k_norm")(k).to(v), v

    def forward(self, img, txt, vec, seg_ids, freqs_cis):
        i_s1, i_c1, i_g1, i_s2, i_c2, i_g2 = self.img_mod(vec).chunk(6, dim=-1)
        t_s1, t_c1, t_g1, t_s2, t_c2, t_g2 = self.txt_mod(vec).chunk(6, dim=-1)
        iq, ik, iv = self._qkv("img", img, i_s1, i_c1)
        if freqs_cis is not None:
            iq, ik = apply_rotary_emb(iq, ik, freqs_cis)
        tq, tk, tv = self._qkv("txt", txt, t_s1, t_c1)
        attn = segment_attention(torch.cat((iq, tq), 1), torch.cat((ik, tk), 1), torch.cat((iv, tv), 1), seg_ids)  # image FIRST
        n = img.shape[1]
        img = img + apply_gate(self.img_attn_proj(attn[:, :n]), i_g1)
        img = img + apply_gate(self.img_mlp(modulate(_ln(img), i_s2, i_c2)), i_g2)
        txt = txt + apply_gate(self.txt_attn_proj(attn[:, n:]), t_g1)
        txt = txt + apply_gate(self.txt_mlp(modulate(_ln(txt), t_s2, t_c2)), t_g2)
        return img, txt


class MMSingleStreamBlock(nn.Module):
    def __init__(self, hidden, heads, mlp_ratio=4):
        super().__init__()
        self.hidden, self.heads, self.mlp_hidden = hidden, heads, mlp_ratio * hidden
        self.linear1, self.linear2 = nn.Linear(hidden, 3 * hidden + self.mlp_hidden), nn.Linear(hidden + self.mlp_hidden, hidden)
        self.q_norm, self.k_norm = RMSNorm(hidden // heads), RMSNorm(hidden // heads)
        self.mlp_act, self.modulation = nn.GELU(approximate="tanh"), ModulateDiT(hidden, 3)

    def forward(self, x, vec, txt_len, seg_ids, freqs_cis):
        shift, scale, gate = self.modulation(vec).chunk(3, dim=-1)
        qkv, mlp = torch.split(self.linear1(modulate(_ln(x), shift, scale)), [3 * self.hidden, self.mlp_hidden], dim=-1)
        b, l = x.shape[:2]
        q, k, v = qkv.view(b, l, 3, self.heads, -1).unbind(2)
        q, k = self.q_norm(q).to(v), self.k_norm(k).to(v)
        if freqs_cis is not None:
            iq, ik = apply_rotary_emb(q[:, :-txt_len], k[:, :-txt_len], freqs_cis)
            q, k = torch.cat((iq, q[:, -txt_len:]), 1), torch.cat((ik, k[:, -txt_len:]), 1)
        attn = segment_attention(q, k, v, seg_ids)
        return x + apply_gate(self.linear2(torch.cat((attn, self.mlp_act(mlp)), 2)), gate)


class IndividualTokenRefinerBlock(nn.Module):
    def __init__(self, hidden, heads, mlp_ratio=4):
        super().__init__()
        self.heads = heads
        self.norm1, self.norm2 = nn.LayerNorm(hidden, eps=1e-6), nn.LayerNorm(hidden, eps=1e-6)
        self.self_attn_qkv, self.self_attn_proj = nn.Linear(hidden, 3 * hidden), nn.Linear(hidden, hidden)
        self.self_attn_q_norm, self.self_attn_k_norm = RMSNorm(hidden // heads), RMSNorm(hidden // heads)
        self.mlp = MLP(hidden, mlp_ratio * hidden, nn.SiLU())
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 2 * hidden))

    def forward(self, x, c, attn_mask):
        gate_msa, gate_mlp = self.adaLN_modulation(c).chunk(2, dim=1)
        b, l = x.shape[:2]
        q, k, v = self.self_attn_qkv(self.norm1(x)).view(b, l, 3, self.heads, -1).unbind(2)
        q, k = self.self_attn_q_norm(q).to(v), self.self_attn_k_norm(k).to(v)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=attn_mask)
        x = x + apply_gate(self.self_attn_proj(o.transpose(1, 2).reshape(b, l, -1)), gate_msa)
        return x + apply_gate(self.mlp(self.norm2(x)), gate_mlp)


class SingleTokenRefiner(nn.Module):
    def __init__(self, in_dim, hidden, heads, depth=2):
        super().__init__()
        self.input_embedder = nn.Linear(in_dim, hidden)
        self.t_embedder, self.c_embedder = TimestepEmbedder(hidden), TextProjection(in_dim, hidden)
        self.individual_token_refiner = nn.Module()
        self.individual_token_refiner.blocks = nn.ModuleList([IndividualTokenRefinerBlock(hidden, heads) for _ in range(depth)])

    def forward(self, x, t, mask):
        mf = mask.to(x.dtype).unsqueeze(-1)
        c = self.t_embedder(t) + self.c_embedder((x * mf).sum(dim=1) / mf.sum(dim=1))
        x = self.input_embedder(x)
        b, l = mask.shape
        m1 = mask.bool().view(b, 1, 1, l).repeat(1, 1, l, 1)
        am = (m1 & m1.transpose(2, 3))
        am[:, :, :, 0] = True
        for blk in self.individual_token_refiner.blocks:
            x = blk(x, c, am)
        return x


class FinalLayer(nn.Module):
    def __init__(self, hidden, out_features):
        super().__init__()
        self.linear = nn.Linear(hidden, out_features)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 2 * hidden))

    def forward(self, x, c):
        shift, scale = self.adaLN_modulation(c).chunk(2, dim=1)  # shift FIRST
        return self.linear(modulate(_ln(x), shift, scale))


class PatchEmbed(nn.Module):
    def __init__(self, patch, in_ch, hidden):
        super().__init__()
        self.proj = nn.Conv3d(in_ch, hidden, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class HYVideoDiffusionTransformer(nn.Module):
    """hyvideo attribute names; HunyuanVideo: hidden 3072, 24 heads, 20 double + 40 single blocks, text 4096 / pooled 768, patch (1,2,2)."""

    def __init__(self, hidden_size=3072, heads_num=24, mm_double_blocks_depth=20, mm_single_blocks_depth=40, in_channels=16,
                 text_states_dim=4096, text_states_dim_2=768, patch_size=(1, 2, 2), guidance_embed=True):
        super().__init__()
        self.patch_size, self.in_channels, self.out_channels = list(patch_size), in_channels, in_channels
        self.hidden_size, self.heads_num, self.guidance_embed = hidden_size, heads_num, guidance_embed
        self.text_projection, self.use_attention_mask = "single_refiner", True
        self.img_in = PatchEmbed(patch_size, in_channels, hidden_size)
        self.txt_in = SingleTokenRefiner(text_states_dim, hidden_size, heads_num, depth=2)
        self.time_in, self.vector_in = TimestepEmbedder(hidden_size), MLPEmbedder(text_states_dim_2, hidden_size)
        if guidance_embed:
            self.guidance_in = TimestepEmbedder(hidden_size)
        self.double_blocks = nn.ModuleList([MMDoubleStreamBlock(hidden_size, heads_num) for _ in range(mm_double_blocks_depth)])
        self.single_blocks = nn.ModuleList([MMSingleStreamBlock(hidden_size, heads_num) for _ in range(mm_single_blocks_depth)])
        self.final_layer = FinalLayer(hidden_size, math.prod(patch_size) * self.out_channels)

    def unpatchify(self, x, t, h, w):
        c, (pt, ph, pw) = self.out_channels, self.patch_size
        x = x.reshape(x.shape[0], t, h, w, c, pt, ph, pw)
        x = torch.einsum("nthwcopq->nctohpwq", x)
        return x.reshape(x.shape[0], c, t * pt, h * ph, w * pw)

    @torch.no_grad()
    def init_synthetic(self, seed=0):
        g = torch.Generator().manual_seed(seed)
        for name, p in self.named_parameters():
            if p.dim() == 1 and ("norm" in name) and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                fan_out, fan_in = p.shape[0], p[0].numel()
                a = math.sqrt(6.0 / (fan_in + fan_out))
                if "mod" in name or "adaLN" in name:
                    a *= 0.3
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * a)
        return self.to(torch.bfloat16)


def seg_ids_from_mask(text_mask, img_len):
    """get_cu_seqlens: per sample the segments [0, img_len + valid) and [img_len + valid, img_len + text_len)."""
    b, l = text_mask.shape
    valid = text_mask.sum(dim=1)
    pos = torch.arange(img_len + l)[None].repeat(b, 1)
    return (pos >= (img_len + valid)[:, None]).long() + 2 * torch.arange(b)[:, None]


def magcache_forward(self, x, t, text_states=None, text_mask=None, text_states_2=None, freqs_cos=None, freqs_sin=None, guidance=None,
                     return_dict=True):
    """MagCache4HunyuanVideo/magcache_sample_video.py:29-160."""
    img, txt = x, text_states
    _, _, ot, oh, ow = x.shape
    tt, th, tw = ot // self.patch_size[0], oh // self.patch_size[1], ow // self.patch_size[2]
    vec = self.time_in(t)                                                  # :52
    vec = vec + self.vector_in(text_states_2)                              # :54
    if self.guidance_embed:
        if guidance is None:
            raise ValueError("Didn't get guidance strength for guidance distilled model.")
        vec = vec + self.guidance_in(guidance)                             # :63
    img = self.img_in(img)                                                 # :65
    txt = self.txt_in(txt, t, text_mask if self.use_attention_mask else None)  # :69
    txt_seq_len, img_seq_len = txt.shape[1], img.shape[1]
    seg_ids = seg_ids_from_mask(text_mask, img_seq_len)                    # :82 (cu_seqlens)
    freqs_cis = (freqs_cos, freqs_sin) if freqs_cos is not None else None
    skip_forward = False
    if self.cnt >= int(self.retention_ratio * self.num_steps):             # :90-102
        cur_mag_ratio = self.mag_ratios[self.cnt]
        self.accumulated_ratio = self.accumulated_ratio * cur_mag_ratio
        cur_skip_err = np.abs(1 - self.accumulated_ratio)
        self.accumulated_err += cur_skip_err
        self.accumulated_steps += 1
        if self.accumulated_err <= self.magcache_thresh and self.accumulated_steps <= self.K:
            cur_residual = self.residual_cache
            skip_forward = True
        else:
            self.accumulated_ratio = 1.0
            self.accumulated_steps = 0
            self.accumulated_err = 0
    if skip_forward:
        img = img + cur_residual                                           # :104
    else:
        ori_img = img
        for block in self.double_blocks:                                   # :108-120
            img, txt = block(img, txt, vec, seg_ids, freqs_cis)
        xx = torch.cat((img, txt), 1)                                      # :123
        for block in self.single_blocks:                                   # :125-137
            xx = block(xx, vec, txt_seq_len, seg_ids, freqs_cis)
        img = xx[:, :img_seq_len, ...]                                     # :139
        cur_residual = img - ori_img                                       # :140
    self.residual_cache = cur_residual                                     # :141
    self.last_skip = skip_forward  # (oracle-only bookkeeping for the tests)
    img = self.final_layer(img, vec)                                       # :144
    img = self.unpatchify(img, tt, th, tw)
    self.cnt += 1                                                          # :149-154
    if self.cnt >= self.num_steps:
        self.cnt = 0
        self.accumulated_ratio = 1.0
        self.accumulated_steps = 0
        self.accumulated_err = 0
    if return_dict:
        return {"x": img}
    return img


def install_magcache(model_cls, mag_ratios, num_steps, thresh=0.24, K=6, retention_ratio=0.2):
    """magcache_sample_video.py:303-328."""
    from .controller_ref import nearest_interp
    model_cls.forward = magcache_forward
    model_cls.cnt, model_cls.num_steps, model_cls.magcache_thresh, model_cls.K = 0, num_steps, thresh, K
    model_cls.residual_cache = None
    mr = np.asarray(mag_ratios, dtype=np.float64)
    if len(mr) != num_steps:
        mr = nearest_interp(mr, num_steps)
    model_cls.mag_ratios, model_cls.retention_ratio = mr, retention_ratio
    model_cls.accumulated_ratio, model_cls.accumulated_err, model_cls.accumulated_steps = 1, 0, 0


def rope_cos_sin(grid, head_dim=128, axes=(16, 56, 56), theta=256.0):
    """get_nd_rotary_pos_embed(rope_dim_list, (t, h, w), theta=256, use_real=True) of the pipeline [EXT]: cos / sin fp32 [t*h*w, 128]."""
    cos, sin = [], []
    mesh = torch.meshgrid(*[torch.arange(n, dtype=torch.float32) for n in grid], indexing="ij")
    for pos, d in zip(mesh, axes):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2)[: d // 2].float() / d))
        ang = torch.outer(pos.reshape(-1), freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1))
        sin.append(ang.sin().repeat_interleave(2, dim=1))
    return torch.cat(cos, dim=1), torch.cat(sin, dim=1)
