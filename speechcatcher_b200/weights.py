"""Pack a reference state_dict into the tensors the CUDA engine consumes.

Key names on the left are the reference's (ESPnet layout, speechcatcher/model/checkpoint_loader.py:117-139
identity mapping); the packed names on the right are what `sc_engine_set_weight` expects.

Re-layouts (all done once at load time, on the host):
  * q/k/v Linear weights of every attention are concatenated along the output dim so that one GEMM
    produces Q|K|V (self-attention) or K|V (cross-attention);
  * conv2 weight [c2][c][kt][kf] -> [c2][(kt,kf,c)] to match the channels-last implicit-GEMM A rows;
  * embed.out weight columns (c2*19+f2) -> (f2*256+c2) to match the [t][f2][c2] conv2 output.
"""
from __future__ import annotations

from typing import Dict

import torch


def sinusoid_table(max_len: int, d: int) -> torch.Tensor:
    """The reference's positional table (model/layers/positional_encoding.py:38-46), same fp32 ops."""
    import math
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def pack_weights(sd: Dict[str, torch.Tensor], enc_layers: int, dec_layers: int, d_model: int = 256
                 ) -> Dict[str, torch.Tensor]:
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    D = d_model
    out: Dict[str, torch.Tensor] = {"pe": sinusoid_table(5000, D)}
    out["enc.conv1.w"] = f32(sd["encoder.embed.conv.0.weight"]).reshape(D, 9)
    out["enc.conv1.b"] = f32(sd["encoder.embed.conv.0.bias"])
    out["enc.conv2.w"] = f32(sd["encoder.embed.conv.2.weight"]).permute(0, 2, 3, 1).reshape(D, 9 * D).contiguous()
    out["enc.conv2.b"] = f32(sd["encoder.embed.conv.2.bias"])
    wo = f32(sd["encoder.embed.out.weight"])                    # [D][c2*F2 + f2]
    f2 = wo.shape[1] // D
    out["enc.out.w"] = wo.reshape(D, D, f2).permute(0, 2, 1).reshape(D, f2 * D).contiguous()
    out["enc.out.b"] = f32(sd["encoder.embed.out.bias"])

    def cat(prefix, names, suffix):
        return torch.cat([f32(sd[f"{prefix}.{n}.{suffix}"]) for n in names], dim=0).contiguous()

    for l in range(enc_layers):
        p, q = f"encoder.encoders.{l}", f"enc.{l}"
        out[f"{q}.qkv.w"] = cat(p + ".self_attn", ("linear_q", "linear_k", "linear_v"), "weight")
        out[f"{q}.qkv.b"] = cat(p + ".self_attn", ("linear_q", "linear_k", "linear_v"), "bias")
        out[f"{q}.o.w"] = f32(sd[p + ".self_attn.linear_out.weight"])
        out[f"{q}.o.b"] = f32(sd[p + ".self_attn.linear_out.bias"])
        out[f"{q}.ff1.w"] = f32(sd[p + ".feed_forward.w_1.weight"])
        out[f"{q}.ff1.b"] = f32(sd[p + ".feed_forward.w_1.bias"])
        out[f"{q}.ff2.w"] = f32(sd[p + ".feed_forward.w_2.weight"])
        out[f"{q}.ff2.b"] = f32(sd[p + ".feed_forward.w_2.bias"])
        for n in ("1", "2"):
            out[f"{q}.ln{n}.w"] = f32(sd[f"{p}.norm{n}.weight"])
            out[f"{q}.ln{n}.b"] = f32(sd[f"{p}.norm{n}.bias"])
    out["enc.after.w"] = f32(sd["encoder.after_norm.weight"])
    out["enc.after.b"] = f32(sd["encoder.after_norm.bias"])
    out["ctc.w"] = f32(sd["ctc.ctc_lo.weight"])
    out["ctc.b"] = f32(sd["ctc.ctc_lo.bias"])
    out["dec.emb"] = f32(sd["decoder.embed.0.weight"])
    for l in range(dec_layers):
        p, q = f"decoder.decoders.{l}", f"dec.{l}"
        out[f"{q}.self_qkv.w"] = cat(p + ".self_attn", ("linear_q", "linear_k", "linear_v"), "weight")
        out[f"{q}.self_qkv.b"] = cat(p + ".self_attn", ("linear_q", "linear_k", "linear_v"), "bias")
        out[f"{q}.self_o.w"] = f32(sd[p + ".self_attn.linear_out.weight"])
        out[f"{q}.self_o.b"] = f32(sd[p + ".self_attn.linear_out.bias"])
        out[f"{q}.src_q.w"] = f32(sd[p + ".src_attn.linear_q.weight"])
        out[f"{q}.src_q.b"] = f32(sd[p + ".src_attn.linear_q.bias"])
        out[f"{q}.src_kv.w"] = cat(p + ".src_attn", ("linear_k", "linear_v"), "weight")
        out[f"{q}.src_kv.b"] = cat(p + ".src_attn", ("linear_k", "linear_v"), "bias")
        out[f"{q}.src_o.w"] = f32(sd[p + ".src_attn.linear_out.weight"])
        out[f"{q}.src_o.b"] = f32(sd[p + ".src_attn.linear_out.bias"])
        out[f"{q}.ff1.w"] = f32(sd[p + ".feed_forward.w_1.weight"])
        out[f"{q}.ff1.b"] = f32(sd[p + ".feed_forward.w_1.bias"])
        out[f"{q}.ff2.w"] = f32(sd[p + ".feed_forward.w_2.weight"])
        out[f"{q}.ff2.b"] = f32(sd[p + ".feed_forward.w_2.bias"])
        for n in ("1", "2", "3"):
            out[f"{q}.ln{n}.w"] = f32(sd[f"{p}.norm{n}.weight"])
            out[f"{q}.ln{n}.b"] = f32(sd[f"{p}.norm{n}.bias"])
    out["dec.after.w"] = f32(sd["decoder.after_norm.weight"])
    out["dec.after.b"] = f32(sd["decoder.after_norm.bias"])
    out["dec.out.w"] = f32(sd["decoder.output_layer.weight"])
    out["dec.out.b"] = f32(sd["decoder.output_layer.bias"])
    return out


def split_f16(w: torch.Tensor) -> torch.Tensor:
    """fp32 weight [N][K] -> the two fp16 planes [2][N][K] of the split-precision tensor-core GEMM
    (csrc/kernels_gemm_x3.cu): hi = fp16(w), lo = fp16((w - hi) * 2^11); w == hi + lo * 2^-11 to 2^-22 relative."""
    w = w.detach().to(torch.float32)
    hi = w.to(torch.float16)
    if not torch.isfinite(hi).all():
        raise ValueError("weight magnitude exceeds the fp16 range of the split-precision GEMM")
    lo = ((w - hi.to(torch.float32)) * 2048.0).to(torch.float16)
    return torch.stack([hi, lo]).contiguous()


# GEMM weights that get a bf16 copy in the tensor-core mode (and fp16 hi / lo planes in the precise tensor-core mode)
def bf16_names(enc_layers: int, dec_layers: int):
    names = ["enc.out.w", "enc.conv2.w", "ctc.w", "dec.out.w"]
    for l in range(enc_layers):
        names += [f"enc.{l}.{n}.w" for n in ("qkv", "o", "ff1", "ff2")]
    for l in range(dec_layers):
        names += [f"dec.{l}.{n}.w" for n in ("self_qkv", "self_o", "src_q", "src_kv", "src_o", "ff1", "ff2")]
    return names
