"""FLUX.1 MMDiT engine: the transformer behind `magcache_forward` of MagCache4FLUX/magcache_flux.py:234-440 on the
same sm_100a kernels as the Wan path — tcgen05 GEMMs (fused bias / GELU / SiLU / bf16 gated-residual epilogues), the tcgen05 flash
attention over the joint text+image sequence, LN+modulate, per-head RMSNorm + RoPE, the K1/K2 cache kernels.

Block arithmetic follows diffusers' `FluxTransformerBlock` / `FluxSingleTransformerBlock` [EXT, not in the reference tree] as restated
in oracle/flux_ref.py; the forward's own statements (embedders, ids, controller, hit / miss, residual, norm_out / proj_out, counter) are
the reference's (file:line cited in `magcache_flux_forward`, magcache_b200/patch.py).

STATUS: written at the end of round 1 without GPU time left to run it — the parity tests (tests/test_flux_forward_gpu.py) are
opt-in (`MC_RUN_UNVALIDATED=1`) until they have passed on a B200. Nothing on the Wan path depends on this module.

HBM layout (S = n_txt + n_img tokens, text rows FIRST — the order of `torch.cat([encoder_hidden_states, hidden_states], dim=1)`, :384;
D = heads*128; everything bf16 like the reference pipeline, which runs without autocast):
  hs   [S, D]    both residual streams; the double-stream blocks work on the two row ranges, the single-stream blocks on all rows
  x0   [n_img, D] x_embedder output (`ori_hidden_states`)         res [n_img, D]  cached residual (`previous_residual`)
  h    [S, D]    LN+modulate output (GEMM A operand)              qk  [S, 2D]     q | k projections, per-head RMSNorm + RoPE in place
  vt   [D, Spad] V^T straight out of the V-projection GEMMs       cat [S, 5D]     single blocks: attention output | GELU(proj_mlp);
                                                                                  double blocks borrow cat[:, D:] as the FF hidden
  ada  [R]       ALL AdaLayerNorm projections of the forward from ONE GEMM over silu(temb) (they depend on temb only)
"""
import numpy as np
import torch

from . import _lib, ops

E = _lib


def _w(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).contiguous()


def _b(t, dev):
    return t.detach().to(device=dev, dtype=torch.bfloat16).float().contiguous()  # bf16 parameter values, kept as fp32 for the epilogues


class FluxWeights:
    """Weights of one FluxTransformer2DModel (diffusers attribute names), repacked: q|k weights concatenated, every AdaLayerNorm
    projection stacked into one matrix, biases / norm weights as fp32 copies of their bf16 values."""

    def __init__(self):
        self.double, self.single = [], []

    @classmethod
    def from_module(cls, m, dev):
        cfg = m.config
        w = cls()
        w.heads, w.head_dim = cfg.num_attention_heads, cfg.attention_head_dim
        if w.head_dim != 128 or tuple(cfg.axes_dims_rope) != (16, 56, 56):
            raise NotImplementedError("FLUX engine: head_dim 128 with RoPE axes (16, 56, 56)")
        w.dim = D = w.heads * w.head_dim
        w.in_channels, w.joint_dim, w.pooled_dim = cfg.in_channels, cfg.joint_attention_dim, cfg.pooled_projection_dim
        w.guidance = bool(cfg.guidance_embeds)
        w.x_w, w.x_b = _w(m.x_embedder.weight, dev), _b(m.x_embedder.bias, dev)
        w.ctx_w, w.ctx_b = _w(m.context_embedder.weight, dev), _b(m.context_embedder.bias, dev)
        tte = m.time_text_embed

        def mlp(e):
            return (_w(e.linear_1.weight, dev), _b(e.linear_1.bias, dev), _w(e.linear_2.weight, dev), _b(e.linear_2.bias, dev))

        w.t_mlp, w.p_mlp = mlp(tte.timestep_embedder), mlp(tte.text_embedder)
        w.g_mlp = mlp(tte.guidance_embedder) if w.guidance else None
        ada_w, ada_b, off = [], [], 0

        def ada(lin):
            nonlocal off
            ada_w.append(lin.weight.detach())
            ada_b.append(lin.bias.detach())
            start, off = off, off + lin.weight.shape[0]
            return start

        for blk in m.transformer_blocks:
            a = blk.attn
            w.double.append({
                "ada": ada(blk.norm1.linear), "ada_c": ada(blk.norm1_context.linear),
                "qk_w": _w(torch.cat([a.to_q.weight, a.to_k.weight], 0), dev), "qk_b": _b(torch.cat([a.to_q.bias, a.to_k.bias], 0), dev),
                "v_w": _w(a.to_v.weight, dev), "v_b": _b(a.to_v.bias, dev), "o_w": _w(a.to_out[0].weight, dev), "o_b": _b(a.to_out[0].bias, dev),
                "nq": _b(a.norm_q.weight, dev), "nk": _b(a.norm_k.weight, dev),
                "cqk_w": _w(torch.cat([a.add_q_proj.weight, a.add_k_proj.weight], 0), dev),
                "cqk_b": _b(torch.cat([a.add_q_proj.bias, a.add_k_proj.bias], 0), dev),
                "cv_w": _w(a.add_v_proj.weight, dev), "cv_b": _b(a.add_v_proj.bias, dev),
                "co_w": _w(a.to_add_out.weight, dev), "co_b": _b(a.to_add_out.bias, dev),
                "cnq": _b(a.norm_added_q.weight, dev), "cnk": _b(a.norm_added_k.weight, dev),
                "ff1_w": _w(blk.ff.net[0].proj.weight, dev), "ff1_b": _b(blk.ff.net[0].proj.bias, dev),
                "ff2_w": _w(blk.ff.net[2].weight, dev), "ff2_b": _b(blk.ff.net[2].bias, dev),
                "cff1_w": _w(blk.ff_context.net[0].proj.weight, dev), "cff1_b": _b(blk.ff_context.net[0].proj.bias, dev),
                "cff2_w": _w(blk.ff_context.net[2].weight, dev), "cff2_b": _b(blk.ff_context.net[2].bias, dev),
            })
        for blk in m.single_transformer_blocks:
            a = blk.attn
            w.single.append({
                "ada": ada(blk.norm.linear),
                "qk_w": _w(torch.cat([a.to_q.weight, a.to_k.weight], 0), dev), "qk_b": _b(torch.cat([a.to_q.bias, a.to_k.bias], 0), dev),
                "v_w": _w(a.to_v.weight, dev), "v_b": _b(a.to_v.bias, dev), "nq": _b(a.norm_q.weight, dev), "nk": _b(a.norm_k.weight, dev),
                "mlp_w": _w(blk.proj_mlp.weight, dev), "mlp_b": _b(blk.proj_mlp.bias, dev),
                "out_w": _w(blk.proj_out.weight, dev), "out_b": _b(blk.proj_out.bias, dev),
            })
        w.ada_out = ada(m.norm_out.linear)
        w.ada_w, w.ada_b, w.ada_rows = _w(torch.cat(ada_w, 0), dev), _b(torch.cat(ada_b, 0), dev), off
        w.out_w, w.out_b = _w(m.proj_out.weight, dev), _b(m.proj_out.bias, dev)
        w.device = dev
        return w


def rope_table(ids, device, axes_dim=(16, 56, 56), theta=10000.0):
    """cos / sin of `FluxPosEmbed` (float64 angles, fp32 values) for ids [S, 3], stored [S, 128] as interleaved (cos, sin) pairs."""
    pos = ids.detach().double().cpu().numpy()
    ang = [np.outer(pos[:, i], 1.0 / theta ** (np.arange(0, d, 2, dtype=np.float64) / d)) for i, d in enumerate(axes_dim)]
    ang = np.concatenate(ang, axis=1)  # [S, 64]
    cs = np.stack([np.cos(ang), np.sin(ang)], axis=-1).reshape(len(pos), 2 * ang.shape[1])
    return torch.from_numpy(cs.astype(np.float32)).to(device)


class FluxEngine:
    def __init__(self, weights: FluxWeights):
        self.w, self.device = weights, weights.device
        self._shape = None
        self._rope_key, self._rope = None, None
        self.res_valid = False

    def _workspace(self, n_img, n_txt):
        if self._shape == (n_img, n_txt):
            return
        D, dev = self.w.dim, self.device
        S = n_img + n_txt
        bf = dict(dtype=torch.bfloat16, device=dev)
        self.n_img, self.n_txt, self.S = n_img, n_txt, S
        self.hs, self.h, self.att = torch.empty(S, D, **bf), torch.empty(S, D, **bf), torch.empty(S, D, **bf)
        self.x0, self.res, self.hit = torch.empty(n_img, D, **bf), torch.empty(n_img, D, **bf), torch.empty(n_img, D, **bf)
        self.qk = torch.empty(S, 2 * D, **bf)
        self.vt = torch.zeros(D, (S + 7) // 8 * 8, **bf)
        # V^T column ranges must start on 16 bytes for the GEMM to write them directly; otherwise (text length not a multiple of 8:
        # never with the pipeline's padded 512 tokens) V goes to a row-major buffer and is transposed once per attention
        self.v_direct = n_txt % 8 == 0
        self.v = None if self.v_direct else torch.empty(S, D, **bf)
        self.cat = torch.empty(S, 5 * D, **bf)
        self.ada = torch.empty(1, self.w.ada_rows, **bf)
        self.adaf = torch.empty(self.w.ada_rows, dtype=torch.float32, device=dev)
        self.s_hidden = torch.empty(n_img, self.w.in_channels, **bf)
        self.s_enc = torch.empty(n_txt, self.w.joint_dim, **bf)
        self.s_pooled = torch.empty(1, self.w.pooled_dim, **bf)
        self.s_t = torch.zeros(2, dtype=torch.float64, device=dev)  # timestep*1000, guidance*1000 (already rounded like the reference)
        self.res_valid = False
        self._shape = (n_img, n_txt)

    # ------------------------------------------------------------------------------------------ inputs (:290-319)
    def stage_inputs(self, hidden_states, encoder_hidden_states, pooled, timestep, guidance, img_ids, txt_ids):
        w = self.w
        assert hidden_states.shape[0] == 1 and encoder_hidden_states.shape[0] == 1, "one sample per call"
        n_img, n_txt = hidden_states.shape[1], encoder_hidden_states.shape[1]
        self._workspace(n_img, n_txt)
        self.s_hidden.copy_(hidden_states[0])
        self.s_enc.copy_(encoder_hidden_states[0])
        self.s_pooled.copy_(pooled.reshape(1, -1))
        # `timestep.to(hidden_states.dtype) * 1000` (:292-294): both the cast and the product round to bf16
        tv = (timestep.reshape(-1)[:1].to(torch.bfloat16) * 1000).double()
        gv = (guidance.reshape(-1)[:1].to(torch.bfloat16) * 1000).double() if guidance is not None else torch.zeros(1, dtype=torch.float64, device=tv.device)
        self.s_t.copy_(torch.cat([tv, gv.to(tv.device)]))
        if (w.guidance and guidance is None) or (not w.guidance and guidance is not None):
            raise ValueError("guidance must be given exactly when the model has guidance_embeds")
        key = (img_ids.data_ptr(), txt_ids.data_ptr(), n_img, n_txt)
        if self._rope_key != key:  # ids are constant over a generation
            self._rope = rope_table(torch.cat((txt_ids.reshape(-1, 3), img_ids.reshape(-1, 3)), dim=0), self.device)  # :318
            self._rope_key = key
            assert self._rope.shape == (self.S, 128)

    def _sinusoid(self, i):
        """`self.time_proj(t).to(dtype=pooled_projection.dtype)`: 256-channel [cos | sin] embedding, rounded to bf16."""
        f = ops.time_sinusoid(self.s_t[i:i + 1], 256)
        return ops.cast_into(f, torch.empty(1, 256, dtype=torch.bfloat16, device=self.device))

    def _time_mlp(self, x_bf16, mlp):
        w1, b1, w2, b2 = mlp
        return ops.gemm(ops.gemm(x_bf16, w1, b1, E.MC_EPI_BIAS_SILU_BF16), w2, b2, E.MC_EPI_BIAS_BF16)

    def prologue(self):
        """x_embedder, time_text_embed, context_embedder (:290-303) and every AdaLayerNorm projection of the forward."""
        w = self.w
        ops.gemm(self.s_hidden, w.x_w, w.x_b, E.MC_EPI_BIAS_BF16, out=self.x0)
        temb = self._time_mlp(self._sinusoid(0), w.t_mlp)
        if w.guidance:
            temb = ops.cache_hit_add(temb, self._time_mlp(self._sinusoid(1), w.g_mlp))
        temb = ops.cache_hit_add(temb, self._time_mlp(self.s_pooled, w.p_mlp))
        ops.gemm(ops.silu(temb), w.ada_w, w.ada_b, E.MC_EPI_BIAS_BF16, out=self.ada)
        ops.cast_into(self.ada.view(-1), self.adaf)
        ops.gemm(self.s_enc, w.ctx_w, w.ctx_b, E.MC_EPI_BIAS_BF16, out=self.hs[:self.n_txt])
        return self.x0

    def _em(self, start, k):
        D = self.w.dim
        return self.adaf[start:start + k * D].view(k, D)

    # ------------------------------------------------------------------------------------------ attention over the joint sequence
    def _attention(self, rows, h_rows, qk_w, qk_b, v_w, v_b, nq, nk):
        """q | k and V^T projections of the token range `rows` (a slice) from its LN output, then per-head RMSNorm + RoPE in place."""
        D, H = self.w.dim, self.w.heads
        ops.gemm(h_rows, qk_w, qk_b, E.MC_EPI_BIAS_BF16, out=self.qk[rows])
        if self.v_direct:
            ops.gemm(v_w, h_rows, v_b, E.MC_EPI_ROWBIAS_BF16, out=self.vt[:, rows])
        else:
            ops.gemm(h_rows, v_w, v_b, E.MC_EPI_BIAS_BF16, out=self.v[rows])
        rope = self._rope[rows]
        ops.rmsnorm_head_rope_(self.qk[rows][:, :D], nq, H, rope)
        ops.rmsnorm_head_rope_(self.qk[rows][:, D:], nk, H, rope)

    def _joint_attention(self, out):
        D, S = self.w.dim, self.S
        if not self.v_direct:
            ops.transpose(self.v, self.vt[:, :S])
        ops.attention(self.qk[:, :D], self.qk[:, D:], self.vt[:, :S], self.w.heads, out=out, tag="flux_attn")

    def run_blocks(self):
        """The double-stream and single-stream blocks (:343-424) on `hs`; returns the image rows of the stream."""
        w, D, H, nt, S = self.w, self.w.dim, self.w.heads, self.n_txt, self.S
        txt, img = slice(0, nt), slice(nt, S)
        hs, h = self.hs, self.h
        hs[img].copy_(self.x0)  # `ori_hidden_states` stays in x0
        for b in w.double:
            em, emc = self._em(b["ada"], 6), self._em(b["ada_c"], 6)
            ops.ln_modulate(hs[img], em, 1, 0, round_ln_to_bf16=True, out=h[img])
            ops.ln_modulate(hs[txt], emc, 1, 0, round_ln_to_bf16=True, out=h[txt])
            self._attention(img, h[img], b["qk_w"], b["qk_b"], b["v_w"], b["v_b"], b["nq"], b["nk"])
            self._attention(txt, h[txt], b["cqk_w"], b["cqk_b"], b["cv_w"], b["cv_b"], b["cnq"], b["cnk"])
            self._joint_attention(self.att)
            ops.gemm(self.att[img], b["o_w"], b["o_b"], E.MC_EPI_BIAS_GATE_RESID_BF16, out=hs[img], gate=em[2])
            ops.gemm(self.att[txt], b["co_w"], b["co_b"], E.MC_EPI_BIAS_GATE_RESID_BF16, out=hs[txt], gate=emc[2])
            for rows, e, f1w, f1b, f2w, f2b in ((img, em, b["ff1_w"], b["ff1_b"], b["ff2_w"], b["ff2_b"]),
                                                (txt, emc, b["cff1_w"], b["cff1_b"], b["cff2_w"], b["cff2_b"])):
                ops.ln_modulate(hs[rows], e, 4, 3, round_ln_to_bf16=True, out=h[rows])
                ffh = self.cat[rows][:, D:]
                ops.gemm(h[rows], f1w, f1b, E.MC_EPI_BIAS_GELU_BF16, out=ffh)
                ops.gemm(ffh, f2w, f2b, E.MC_EPI_BIAS_GATE_RESID_BF16, out=hs[rows], gate=e[5])
        allr = slice(0, S)
        for b in w.single:
            em = self._em(b["ada"], 3)
            ops.ln_modulate(hs, em, 1, 0, round_ln_to_bf16=True, out=h)
            ops.gemm(h, b["mlp_w"], b["mlp_b"], E.MC_EPI_BIAS_GELU_BF16, out=self.cat[:, D:])
            self._attention(allr, h, b["qk_w"], b["qk_b"], b["v_w"], b["v_b"], b["nq"], b["nk"])
            self._joint_attention(self.cat[:, :D])
            ops.gemm(self.cat, b["out_w"], b["out_b"], E.MC_EPI_BIAS_GATE_RESID_BF16, out=hs, gate=em[2])
        return hs[img]

    def head(self, x_img):
        """`norm_out(hidden_states, temb)`, `proj_out` (:429-430): AdaLayerNormContinuous chunks (scale, shift) in that order."""
        w = self.w
        em = self._em(w.ada_out, 2)
        ops.ln_modulate(x_img, em, 0, 1, round_ln_to_bf16=True, out=self.h[self.n_txt:])
        return ops.gemm(self.h[self.n_txt:], w.out_w, w.out_b, E.MC_EPI_BIAS_BF16)

    def forward(self, kind):
        x0 = self.prologue()
        if kind == "hit":
            if not self.res_valid:
                raise TypeError("magcache_b200: cache hit with an empty previous_residual (reference: Tensor + NoneType)")
            x = ops.cache_hit_add(x0, self.res, out=self.hit)                     # :340
        else:
            x = self.run_blocks()
            ops.residual_sub(x.contiguous(), x0, out=self.res)                    # :426 (x is a contiguous row range of hs)
            self.res_valid = True
        return self.head(x)
