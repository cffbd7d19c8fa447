"""ORACLE — test infrastructure only (imported by tests/ and __graft_entry__.smoke(), never by the product).

CPU restatement of the FLUX path of the reference: `magcache_forward` of MagCache4FLUX/magcache_flux.py:234-440 (the patched
`FluxTransformer2DModel.forward`) and the transformer it drives.

PARITY UNPINNED for the block arithmetic: the modules (`FluxTransformerBlock`, `FluxSingleTransformerBlock`, `AdaLayerNormZero*`,
`FluxAttnProcessor2_0`, `FluxPosEmbed`, `CombinedTimestep(Guidance)TextProjEmbeddings`, `RMSNorm`, `FeedForward`) live in `diffusers`,
which is not under /root/reference and is effectively unpinned by it; what follows restates them from SURVEY Appendix B.2
under the attribute names diffusers uses, so that `FluxWeights.from_module` reads a real pipeline's transformer the same way.
The controller / cache / counter statements ARE the reference's (file:line cited inline) and are pinned by tests/golden/masks.json.

dtype rule of the FLUX scripts: the pipeline is loaded in bf16 and runs WITHOUT autocast, so every tensor — streams, modulation
vectors, the residual cache — is bf16 and every torch op rounds to bf16 (`DT`). `exact()` switches the same code to float64.
"""
import math
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

DT = torch.bfloat16


class exact:
    """Evaluate the same network with no reduced-precision rounding (tests: the point both implementations are measured against)."""

    def __enter__(self):
        global DT
        self._old, DT = DT, torch.float64

    def __exit__(self, *a):
        global DT
        DT = self._old


# ----------------------------------------------------------------------------------------------- embeddings [EXT diffusers]
def get_timestep_embedding(timesteps, dim=256, max_period=10000):
    """Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0): fp32 sinusoid, [cos | sin]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, dim):
        super().__init__()
        self.linear_1, self.act, self.linear_2 = nn.Linear(in_channels, dim), nn.SiLU(), nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    def __init__(self, dim, pooled_dim, guidance_embeds=True):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim)
        if guidance_embeds:
            self.guidance_embedder = TimestepEmbedding(256, dim)
        self.text_embedder = TimestepEmbedding(pooled_dim, dim)  # PixArtAlphaTextProjection(act_fn="silu"): linear_1, silu, linear_2

    def forward(self, timestep, guidance, pooled_projection):
        dt = pooled_projection.dtype
        cond = self.timestep_embedder(get_timestep_embedding(timestep).to(dt))
        if guidance is not None:
            cond = cond + self.guidance_embedder(get_timestep_embedding(guidance).to(dt))
        return cond + self.text_embedder(pooled_projection)


def rope_freqs(ids, axes_dim=(16, 56, 56), theta=10000.0):
    """FluxPosEmbed: per axis `get_1d_rotary_pos_embed(dim_i, ids[:, i], repeat_interleave_real=True, use_real=True,
    freqs_dtype=float64)`; returns (cos, sin) fp32 [S, sum(axes_dim)], every frequency repeated for its (real, imag) pair."""
    cos, sin = [], []
    pos = ids.to(torch.float64)
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64) / d))
        ang = torch.outer(pos[:, i], freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos, dim=-1), torch.cat(sin, dim=-1)


def apply_rotary_emb(x, freqs_cis):
    """x [B, H, S, D]; consecutive (real, imag) pairs; fp32 arithmetic, result in x's dtype."""
    cos, sin = freqs_cis
    cos, sin = cos[None, None].to(torch.float64 if DT == torch.float64 else torch.float32), sin[None, None].to(torch.float64 if DT == torch.float64 else torch.float32)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    wide = torch.float64 if DT == torch.float64 else torch.float32
    return (x.to(wide) * cos + x_rot.to(wide) * sin).to(x.dtype)


# ----------------------------------------------------------------------------------------------- norms [EXT diffusers]
class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps, self.weight = eps, nn.Parameter(torch.ones(dim))

    def forward(self, h):
        wide = torch.float64 if DT == torch.float64 else torch.float32
        variance = h.to(wide).pow(2).mean(-1, keepdim=True)
        h = h * torch.rsqrt(variance + self.eps)  # bf16 * fp32 -> fp32
        return h.to(self.weight.dtype) * self.weight


def _ln(x, eps=1e-6):
    return F.layer_norm(x, (x.shape[-1],), None, None, eps)  # bf16 in, fp32 inside, bf16 out


class AdaLayerNormZero(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.silu, self.linear = nn.SiLU(), nn.Linear(dim, 6 * dim)

    def forward(self, x, emb):
        emb = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        return _ln(x) * (1 + scale_msa[:, None]) + shift_msa[:, None], gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormZeroSingle(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.silu, self.linear = nn.SiLU(), nn.Linear(dim, 3 * dim)

    def forward(self, x, emb):
        shift_msa, scale_msa, gate_msa = self.linear(self.silu(emb)).chunk(3, dim=1)
        return _ln(x) * (1 + scale_msa[:, None]) + shift_msa[:, None], gate_msa


class AdaLayerNormContinuous(nn.Module):
    def __init__(self, dim, cond_dim):
        super().__init__()
        self.silu, self.linear = nn.SiLU(), nn.Linear(cond_dim, 2 * dim)

    def forward(self, x, cond):
        scale, shift = self.linear(self.silu(cond).to(x.dtype)).chunk(2, dim=1)  # scale FIRST
        return _ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]


# ----------------------------------------------------------------------------------------------- attention + blocks [EXT diffusers]
class Attention(nn.Module):
    """diffusers `Attention` as FLUX configures it (qk_norm="rms_norm", bias=True) with FluxAttnProcessor2_0 inlined."""

    def __init__(self, dim, heads, head_dim, added_kv=False, pre_only=False):
        super().__init__()
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.norm_q, self.norm_k = RMSNorm(head_dim), RMSNorm(head_dim)
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        if added_kv:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
            self.norm_added_q, self.norm_added_k = RMSNorm(head_dim), RMSNorm(head_dim)
            self.to_add_out = nn.Linear(dim, dim)

    def forward(self, hidden_states, encoder_hidden_states=None, image_rotary_emb=None):
        b, h = hidden_states.shape[0], self.heads

        def split(t):
            return t.view(b, -1, h, t.shape[-1] // h).transpose(1, 2)

        q, k, v = split(self.to_q(hidden_states)), split(self.to_k(hidden_states)), split(self.to_v(hidden_states))
        q, k = self.norm_q(q), self.norm_k(k)
        if encoder_hidden_states is not None:
            eq, ek, ev = (split(p(encoder_hidden_states)) for p in (self.add_q_proj, self.add_k_proj, self.add_v_proj))
            eq, ek = self.norm_added_q(eq), self.norm_added_k(ek)
            q, k, v = torch.cat([eq, q], dim=2), torch.cat([ek, k], dim=2), torch.cat([ev, v], dim=2)  # text tokens FIRST
        if image_rotary_emb is not None:
            q, k = apply_rotary_emb(q, image_rotary_emb), apply_rotary_emb(k, image_rotary_emb)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, -1, h * q.shape[-1]).to(q.dtype)
        if encoder_hidden_states is not None:
            n = encoder_hidden_states.shape[1]
            return self.to_out[0](o[:, n:]), self.to_add_out(o[:, :n])
        return o


class FeedForward(nn.Module):
    """FeedForward(dim, dim_out=dim, activation_fn="gelu-approximate"): net[0] = GELU(proj + tanh-gelu), net[1] dropout, net[2] Linear."""

    def __init__(self, dim, mult=4):
        super().__init__()
        gelu = nn.Module()
        gelu.proj = nn.Linear(dim, dim * mult)
        self.net = nn.ModuleList([gelu, nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        return self.net[2](F.gelu(self.net[0].proj(x), approximate="tanh"))


class FluxTransformerBlock(nn.Module):
    def __init__(self, dim, heads, head_dim):
        super().__init__()
        self.norm1, self.norm1_context = AdaLayerNormZero(dim), AdaLayerNormZero(dim)
        self.attn = Attention(dim, heads, head_dim, added_kv=True)
        self.ff, self.ff_context = FeedForward(dim), FeedForward(dim)

    def forward(self, hidden_states, encoder_hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
        n, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.norm1(hidden_states, temb)
        nc, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = self.norm1_context(encoder_hidden_states, temb)
        attn_output, context_attn_output = self.attn(n, nc, image_rotary_emb)
        hidden_states = hidden_states + gate_msa.unsqueeze(1) * attn_output
        n2 = _ln(hidden_states) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        hidden_states = hidden_states + gate_mlp.unsqueeze(1) * self.ff(n2)
        encoder_hidden_states = encoder_hidden_states + c_gate_msa.unsqueeze(1) * context_attn_output
        nc2 = _ln(encoder_hidden_states) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        encoder_hidden_states = encoder_hidden_states + c_gate_mlp.unsqueeze(1) * self.ff_context(nc2)
        return encoder_hidden_states, hidden_states


class FluxSingleTransformerBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, mlp_ratio=4):
        super().__init__()
        self.norm = AdaLayerNormZeroSingle(dim)
        self.proj_mlp, self.act_mlp = nn.Linear(dim, dim * mlp_ratio), nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(dim + dim * mlp_ratio, dim)
        self.attn = Attention(dim, heads, head_dim, pre_only=True)

    def forward(self, hidden_states, temb, image_rotary_emb=None, joint_attention_kwargs=None):
        residual = hidden_states
        n, gate = self.norm(hidden_states, temb)
        mlp = self.act_mlp(self.proj_mlp(n))
        attn_output = self.attn(n, image_rotary_emb=image_rotary_emb)
        hidden_states = gate.unsqueeze(1) * self.proj_out(torch.cat([attn_output, mlp], dim=2))
        return residual + hidden_states


class FluxTransformer2DModel(nn.Module):
    """Attribute names of diffusers' FluxTransformer2DModel (FLUX.1-dev: 19 + 38 blocks, 24 heads x 128, joint dim 4096, pooled 768)."""

    def __init__(self, in_channels=64, num_layers=19, num_single_layers=38, attention_head_dim=128, num_attention_heads=24,
                 joint_attention_dim=4096, pooled_projection_dim=768, guidance_embeds=True, axes_dims_rope=(16, 56, 56)):
        super().__init__()
        dim = attention_head_dim * num_attention_heads
        self.inner_dim, self.axes_dims_rope, self.guidance_embeds = dim, tuple(axes_dims_rope), guidance_embeds
        self.config = types.SimpleNamespace(in_channels=in_channels, num_layers=num_layers, num_single_layers=num_single_layers,
                                            attention_head_dim=attention_head_dim, num_attention_heads=num_attention_heads,
                                            joint_attention_dim=joint_attention_dim, pooled_projection_dim=pooled_projection_dim,
                                            guidance_embeds=guidance_embeds, axes_dims_rope=tuple(axes_dims_rope))
        self.time_text_embed = CombinedTimestepGuidanceTextProjEmbeddings(dim, pooled_projection_dim, guidance_embeds)
        self.context_embedder, self.x_embedder = nn.Linear(joint_attention_dim, dim), nn.Linear(in_channels, dim)
        self.transformer_blocks = nn.ModuleList([FluxTransformerBlock(dim, num_attention_heads, attention_head_dim) for _ in range(num_layers)])
        self.single_transformer_blocks = nn.ModuleList([FluxSingleTransformerBlock(dim, num_attention_heads, attention_head_dim)
                                                        for _ in range(num_single_layers)])
        self.norm_out, self.proj_out = AdaLayerNormContinuous(dim, dim), nn.Linear(dim, in_channels)

    def pos_embed(self, ids):
        return rope_freqs(ids, self.axes_dims_rope)

    @torch.no_grad()
    def init_synthetic(self, seed=0):
        """Seeded synthetic weights: xavier-uniform matrices (AdaLN projections scaled down so the modulation stays O(0.1)), non-zero
        biases and norm weights; stored in bf16 like a loaded FLUX checkpoint."""
        g = torch.Generator().manual_seed(seed)
        for name, p in self.named_parameters():
            if p.dim() == 1 and "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            else:
                a = math.sqrt(6.0 / (p.shape[0] + p.shape[1]))
                if ".linear.weight" in name and ("norm" in name):  # AdaLN projections
                    a *= 0.3
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * a)
        return self.to(torch.bfloat16)


# ----------------------------------------------------------------------------------------------- the patched forward
def magcache_forward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None, img_ids=None, txt_ids=None,
                     guidance=None, joint_attention_kwargs=None, controlnet_block_samples=None, controlnet_single_block_samples=None,
                     return_dict=True, controlnet_blocks_repeat=False):
    """MagCache4FLUX/magcache_flux.py:234-440 without the LoRA-scale, ip-adapter and ControlNet side paths (None in the reference's
    own script, :473-490)."""
    hidden_states = self.x_embedder(hidden_states)                                     # :290
    timestep = timestep.to(hidden_states.dtype) * 1000                                 # :292
    guidance = guidance.to(hidden_states.dtype) * 1000 if guidance is not None else None
    temb = self.time_text_embed(timestep, guidance, pooled_projections)                # :298-302
    encoder_hidden_states = self.context_embedder(encoder_hidden_states)               # :303
    ids = torch.cat((txt_ids, img_ids), dim=0)                                         # :318
    image_rotary_emb = self.pos_embed(ids)
    skip_forward = False
    if self.cnt >= int(self.retention_ratio * self.num_steps + 0.5):                   # :327-338
        cur_scale = self.mag_ratios[self.cnt]
        self.accumulated_ratio = self.accumulated_ratio * cur_scale
        self.accumulated_steps += 1
        self.accumulated_err += np.abs(1 - self.accumulated_ratio)
        not_step_11 = np.round(self.cnt * ((28 - 1) / (self.num_steps - 1))).astype(int) != 11
        if self.accumulated_err <= self.magcache_thresh and self.accumulated_steps <= self.K and not_step_11:
            cur_residual = self.previous_residual
            skip_forward = True
        else:
            self.accumulated_ratio = 1.0
            self.accumulated_steps = 0
            self.accumulated_err = 0
    if skip_forward:
        hidden_states = hidden_states + cur_residual                                   # :340
    else:
        ori_hidden_states = hidden_states
        for block in self.transformer_blocks:                                          # :343-383
            encoder_hidden_states, hidden_states = block(hidden_states=hidden_states, encoder_hidden_states=encoder_hidden_states,
                                                         temb=temb, image_rotary_emb=image_rotary_emb)
        hidden_states = torch.cat([encoder_hidden_states, hidden_states], dim=1)       # :384
        for block in self.single_transformer_blocks:                                   # :386-422
            hidden_states = block(hidden_states=hidden_states, temb=temb, image_rotary_emb=image_rotary_emb)
        hidden_states = hidden_states[:, encoder_hidden_states.shape[1]:, ...]         # :424
        cur_residual = hidden_states - ori_hidden_states                               # :426
    self.previous_residual = cur_residual                                              # :427
    self.last_skip = skip_forward  # (oracle-only bookkeeping for the tests)
    hidden_states = self.norm_out(hidden_states, temb)                                 # :429-430
    output = self.proj_out(hidden_states)
    self.cnt += 1                                                                      # :431-436
    if self.cnt >= self.num_steps:
        self.cnt = 0
        self.accumulated_ratio = 1.0
        self.accumulated_steps = 0
        self.accumulated_err = 0
    if not return_dict:
        return (output,)
    return types.SimpleNamespace(sample=output)


def magcache_calibration(self, hidden_states, encoder_hidden_states=None, pooled_projections=None, timestep=None, img_ids=None, txt_ids=None,
                         guidance=None, joint_attention_kwargs=None, controlnet_block_samples=None, controlnet_single_block_samples=None,
                         return_dict=True, controlnet_blocks_repeat=False):
    """MagCache4FLUX/magcache_flux.py:21-231 (same omissions as magcache_forward); the statistics are bf16 tensor ops like upstream."""
    hidden_states = self.x_embedder(hidden_states)
    timestep = timestep.to(hidden_states.dtype) * 1000
    guidance = guidance.to(hidden_states.dtype) * 1000 if guidance is not None else None
    temb = self.time_text_embed(timestep, guidance, pooled_projections)
    encoder_hidden_states = self.context_embedder(encoder_hidden_states)
    image_rotary_emb = self.pos_embed(torch.cat((txt_ids, img_ids), dim=0))
    ori_hidden_states = hidden_states
    for block in self.transformer_blocks:
        encoder_hidden_states, hidden_states = block(hidden_states=hidden_states, encoder_hidden_states=encoder_hidden_states, temb=temb,
                                                     image_rotary_emb=image_rotary_emb)
    hidden_states = torch.cat([encoder_hidden_states, hidden_states], dim=1)
    for block in self.single_transformer_blocks:
        hidden_states = block(hidden_states=hidden_states, temb=temb, image_rotary_emb=image_rotary_emb)
    hidden_states = hidden_states[:, encoder_hidden_states.shape[1]:, ...]
    cur_residual = hidden_states - ori_hidden_states                                   # :197
    if self.cnt >= 1:                                                                  # :198-205
        norm_ratio = ((cur_residual.norm(dim=-1) / self.previous_residual.norm(dim=-1)).mean()).item()
        norm_std = (cur_residual.norm(dim=-1) / self.previous_residual.norm(dim=-1)).std().item()
        cos_dis = (1 - F.cosine_similarity(cur_residual, self.previous_residual, dim=-1, eps=1e-8)).mean().item()
        self.norm_ratio.append(round(norm_ratio, 5))
        self.norm_std.append(round(norm_std, 5))
        self.cos_dis.append(round(cos_dis, 5))
    self.previous_residual = cur_residual
    output = self.proj_out(self.norm_out(hidden_states, temb))
    self.cnt += 1
    if self.cnt >= self.num_steps:                                                     # :217-221
        self.cnt = 0
        self.norm_ratio, self.norm_std, self.cos_dis = [], [], []
    return types.SimpleNamespace(sample=output) if return_dict else (output,)


def install_magcache(model_cls, mag_ratios, num_steps, thresh=0.24, K=5, retention_ratio=0.1):
    """MagCache4FLUX/magcache_flux.py:446-471."""
    from .controller_ref import nearest_interp
    model_cls.forward = magcache_forward
    model_cls.cnt, model_cls.num_steps = 0, num_steps
    mr = np.asarray(mag_ratios, dtype=np.float64)
    if len(mr) != num_steps:
        mr = nearest_interp(mr, num_steps)
    model_cls.mag_ratios = mr
    model_cls.K, model_cls.magcache_thresh, model_cls.retention_ratio = K, thresh, retention_ratio
    model_cls.accumulated_ratio, model_cls.accumulated_err, model_cls.accumulated_steps = 1, 0, 0
    model_cls.previous_residual = None


def make_ids(h_tokens, w_tokens, n_text):
    """`_prepare_latent_image_ids` of the FLUX pipeline [EXT]: img ids (0, row, col); text ids all zero."""
    img = torch.zeros(h_tokens, w_tokens, 3)
    img[..., 1] += torch.arange(h_tokens)[:, None]
    img[..., 2] += torch.arange(w_tokens)[None, :]
    return img.reshape(-1, 3), torch.zeros(n_text, 3)
