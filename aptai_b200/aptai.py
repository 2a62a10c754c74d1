"""Drop-in for the reference's models/aptai.py (class APTAI) on the aptai_b200 kernels.

Same constructor, same `forward` / `get_aptai_output` / `get_config` signatures, return keys and state_dict layout
(`wav2vec2.*`, `tv_head.2.*`, `phn_head.2.*`, `tv_lowpass.lowpass.weight`).  The reference hard-codes a 24x1024
backbone (`hidden_states[24]`, `nn.Linear(1024, ...)`, models/aptai.py:46,54,81,141); here both come from the
config (`hidden_states[num_hidden_layers]`, `hidden_size`), which is identical for the 24x1024 case.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .backbone import Wav2Vec2Backbone
from .modules import LowPassFilterLayer
from .graphs import GraphCache
from .train import GradBuffer, GradReducer, attach_backward, broadcast_parameters

TV_NAMES = ("LA", "LP", "JA", "TTCL", "TTCD", "TMCL", "TMCD", "TBCL", "TBCD")


class APTAI(nn.Module):
    def __init__(self, device, vocab, huggingface_model_id, pretrain_cfg, cache_dir, phn_drop=0.1, tv_drop=0.1,
                 freeze_feature_encoder=True):
        super().__init__()
        self.device = device
        self.vocab = vocab
        self.huggingface_model_id = huggingface_model_id
        self.pretrain_cfg = pretrain_cfg
        self.wav2vec2 = Wav2Vec2Backbone.from_pretrained(huggingface_model_id, config=pretrain_cfg,
                                                         cache_dir=cache_dir).to(self.device)
        self.wav2vec2.gradient_checkpointing_enable()
        if freeze_feature_encoder:
            self.wav2vec2.freeze_feature_encoder()
        H = self.wav2vec2.cfg.hidden_size
        # parameter containers with the reference's Sequential indices (Dropout, activation, Linear)
        self.tv_head = nn.Sequential(nn.Dropout(tv_drop), nn.Tanh(), nn.Linear(H, 9))
        self.tv_lowpass = LowPassFilterLayer(self.device, 10, 49, 9)
        self.phn_head = nn.Sequential(nn.Dropout(phn_drop), nn.LeakyReLU(), nn.Linear(H, 46))

    # ---------------------------------------------------------------------------------------------- kernels
    @torch.no_grad()
    def _heads(self, audio_inputs, audio_lengths, want_logp: bool = False):
        """backbone -> ONE tail kernel (final LayerNorm + TV / phoneme heads + argmax [+ log-softmax]) -> low-pass.
        Returns tv [B,T,9], logits [B,T,46], pred [B,T] (and log_probs [B,T,46] with `want_logp`)."""
        if self.training and (self.tv_head[0].p > 0 or self.phn_head[0].p > 0):
            raise RuntimeError("aptai_b200: this is the inference path (no head dropout); in training mode call "
                               "forward() with autograd enabled, or .eval() first")
        w2v = self.wav2vec2
        if not audio_inputs.is_cuda:
            raise RuntimeError("aptai_b200: audio_inputs must be on a CUDA (sm_100) device; there is no CPU path")
        wav = audio_inputs.to(torch.float32).contiguous()
        flen = w2v.frame_lengths_i32(audio_lengths)
        defer = w2v.defers_final_ln()
        h, _, _ = w2v.encode(wav, flen, final_ln=not defer)
        B, T, H = h.shape
        g, b, eps = w2v.final_ln_params() if defer else (None, None, 0.0)
        tvl, phl = self.tv_head[2], self.phn_head[2]
        f = lambda p: p.detach().float().contiguous()
        tv_raw, logits, pred, logp, _ = ops.tail(h.view(B * T, H), g, b, eps, f(tvl.weight), f(tvl.bias), ops.ACT_TANH,
                                                 f(phl.weight), f(phl.bias), ops.ACT_LEAKY, want_logp=want_logp)
        tv = self.tv_lowpass(tv_raw.view(B, T, 9))
        if want_logp:
            return tv, logits.view(B, T, -1), pred.view(B, T), logp.view(B, T, -1), flen
        return tv, logits.view(B, T, -1), pred.view(B, T)

    # ---------------------------------------------------------------------------------------------- training
    def enable_data_parallel(self, group=None, layers_per_bucket: int = 4, broadcast: bool = True,
                             overlap_optimizer: bool = False):
        """Data-parallel training over `torch.distributed` (one process per GPU): weights are broadcast from rank 0
        and every backward all-reduces (averages) the flat gradient buffer, bucketed by encoder layers and
        overlapped with the remaining backward kernels (BASELINE config 4; the reference trains single-GPU).
        `overlap_optimizer=True` (with `FusedAdam`): the backward returns with the last reductions still in flight and
        the optimizer updates each bucket as soon as its all-reduce has landed; code that reads `.grad` between
        `backward()` and `step()` must call `model.grad_buffer().wait_pending()` first."""
        if broadcast:
            broadcast_parameters(self, 0, group)
        gb = self.grad_buffer()
        red = GradReducer(gb, "wav2vec2.encoder.layers.", len(self.wav2vec2.encoder.layers), layers_per_bucket, group,
                          defer_wait=overlap_optimizer)
        object.__setattr__(self, "_reducer", red)
        return red

    def grad_buffer(self) -> GradBuffer:
        """Flat fp32 gradient buffer; (re)attaches `p.grad` views when an optimizer's zero_grad(set_to_none=True)
        dropped them (which also means: start from zero)."""
        gb = getattr(self, "_grad_buffer", None)
        if gb is None:
            gb = GradBuffer(list(self.named_parameters()), self.wav2vec2.fused_grad_groups("wav2vec2."))
            object.__setattr__(self, "_grad_buffer", gb)
            return gb
        dropped = False
        for n, p in gb.params:
            if not gb.owns(p):
                dropped = True
                p.grad = gb.flat[gb.offsets[n]: gb.offsets[n] + p.numel()].view(p.shape)
        if dropped:
            gb.zero()
        return gb

    def _forward_train(self, audio_inputs, audio_lengths, phn_targets, tv_targets):
        """Training step forward: same kernels, activations kept; `loss.backward()` launches the backward kernels.
        Head dropouts (models/aptai.py:44,52) are counter-based like the backbone's (regenerated in the backward)."""
        w2v = self.wav2vec2
        dev = next(w2v.parameters()).device
        wav = audio_inputs.to(device=dev, dtype=torch.float32).contiguous()
        lens = audio_lengths.reshape(-1).to(device=dev, dtype=torch.int64)
        flen = w2v.frame_lengths_i32(lens)
        gb = self.grad_buffer()
        last, sv = w2v.encode_train(wav, flen)
        B, T, H = last.shape
        h = last.view(B * T, H)
        p_tv, p_phn = float(self.tv_head[0].p), float(self.phn_head[0].p)
        s_tv, s_phn = w2v.drop_seed(sv.step, -1, w2v.SITE_HEAD_A), w2v.drop_seed(sv.step, -1, w2v.SITE_HEAD_B)
        h_tv = ops.dropout(h, p_tv, s_tv, want_f32=True)[0] if p_tv > 0 else h
        h_phn = ops.dropout(h, p_phn, s_phn, want_f32=True)[0] if p_phn > 0 else h
        tvl, phl = self.tv_head[2], self.phn_head[2]
        f = lambda p: p.detach().float().contiguous()
        wa, ba, wb, bb = f(tvl.weight), f(tvl.bias), f(phl.weight), f(phl.bias)
        if h_tv is h_phn:
            tv_raw, logits, pred = ops.heads(h, wa, ba, ops.ACT_TANH, wb, bb, ops.ACT_LEAKY)
        else:
            tv_raw, _, _ = ops.heads(h_tv, wa, ba, ops.ACT_TANH, None, None, 0, want_argmax=False)
            _, logits, pred = ops.heads(h_phn, None, None, 0, wb, bb, ops.ACT_LEAKY)
        tv = self.tv_lowpass(tv_raw.view(B, T, 9))
        taps = self.tv_lowpass.lowpass.weight.detach().reshape(-1).contiguous()
        tvt = tv_targets.view(B * T, 9)
        res, ws = ops.masked_mse_ce(tv.view(B * T, 9), tvt, logits, phn_targets, return_ws=True)

        def run_backward(grad_out):
            gs = grad_out.detach().reshape(1).to(device=dev, dtype=torch.float32)
            d_tvlp, d_lg = ops.masked_mse_ce_bwd(tv.view(B * T, 9), tvt, logits, phn_targets, ws, gs)
            d_tv = ops.lowpass(d_tvlp.view(B, T, 9), taps).view(B * T, 9)   # symmetric FIR: adjoint = the filter
            if h_tv is h_phn:
                dh = ops.heads_bwd(h, d_tv, wa, ops.ACT_TANH, gb.view("tv_head.2.weight"), gb.view("tv_head.2.bias"),
                                   d_lg, wb, ops.ACT_LEAKY, gb.view("phn_head.2.weight"), gb.view("phn_head.2.bias"))
            else:
                dh_a = ops.heads_bwd(h_tv, d_tv, wa, ops.ACT_TANH, gb.view("tv_head.2.weight"),
                                     gb.view("tv_head.2.bias"), None, None, 0, None, None)
                dh_b = ops.heads_bwd(h_phn, None, None, 0, None, None, d_lg, wb, ops.ACT_LEAKY,
                                     gb.view("phn_head.2.weight"), gb.view("phn_head.2.bias"))
                if p_tv > 0:
                    ops.dropout(dh_a, p_tv, s_tv, out_f32=dh_a)
                dh = ops.dropout(dh_b, p_phn, s_phn, residual=dh_a, out_f32=dh_b)[0]      # p = 0: a plain add
            red = getattr(self, "_reducer", None)
            w2v.backward(sv, dh, gb, prefix="wav2vec2.", on_layer_done=red.layer_done if red else None)
            if red is not None:
                red.finish()

        loss = attach_backward(res[0], tvl.weight, run_backward)
        return {"loss": loss, "mse_loss": res[1], "ce_loss": res[2], "tvs_pred": tv, "phn_fc_pred": pred.view(B, T)}

    def forward(self, epoch, audio_inputs, audio_lengths, phn_frames_49hz, LA, LP, JA, TTCL, TTCD, TMCL, TMCD, TBCL,
                TBCD):
        """models/aptai.py:58-115.  Loss values come from the fused masked-MSE/CE kernel.  In training mode with
        autograd enabled the returned loss carries the hand-written backward (train/train_aptai.py:440)."""
        tv_targets = torch.stack([LA, LP, JA, TTCL, TTCD, TMCL, TMCD, TBCL, TBCD], dim=-1).float().contiguous()
        if self.training and torch.is_grad_enabled():
            dev = next(self.wav2vec2.parameters()).device
            return self._forward_train(audio_inputs, audio_lengths,
                                       phn_frames_49hz.reshape(-1).to(device=dev, dtype=torch.int64).contiguous(),
                                       tv_targets.to(dev))
        tv, logits, pred = self._heads(audio_inputs, audio_lengths)
        B, T, V = logits.shape
        res = ops.masked_mse_ce(tv.view(B * T, 9), tv_targets.view(B * T, 9).to(tv.device),
                                logits.view(B * T, V), phn_frames_49hz.reshape(-1).to(device=tv.device,
                                                                                     dtype=torch.int64).contiguous())
        return {"loss": res[0], "mse_loss": res[1], "ce_loss": res[2], "tvs_pred": tv, "phn_fc_pred": pred}

    @torch.no_grad()
    def predict(self, audio_inputs, audio_lengths, phn_targets=None, phn_target_lens=None, blank=0):
        """Batched inference (additive API; the reference only offers the single-utterance `get_aptai_output`).
        Returns tvs_pred [B,T,9], phn_fc_logits [B,T,46], phn_fc_pred [B,T]; with known phoneme sequences
        (`phn_targets` int32 [B,S], `phn_target_lens` int32 [B]) also the CTC-Viterbi forced alignment of the
        phoneme log-probs: align_paths int32 [B,T] (-1 beyond each utterance), align_scores, align_status."""
        if phn_targets is None:
            tv, logits, pred = self._heads(audio_inputs, audio_lengths)
            return {"tvs_pred": tv, "phn_fc_logits": logits, "phn_fc_pred": pred}
        tv, logits, pred, lp, flen = self._heads(audio_inputs, audio_lengths, want_logp=True)
        out = {"tvs_pred": tv, "phn_fc_logits": logits, "phn_fc_pred": pred}
        paths, scores, status = ops.ctc_viterbi(lp, phn_targets, flen, phn_target_lens, blank=blank)
        out.update(align_paths=paths, align_scores=scores, align_status=status, log_probs=lp)
        return out

    def set_precision(self, precision: str):
        """'bf16' (default), 'fp16' (same kernels on IEEE fp16 operands: same speed, 7x smaller output error;
        inference only) or 'f32x3' (accuracy mode of the encoder, aptai_b200/accurate.py; inference only)."""
        self.wav2vec2.set_precision(precision)
        return self

    def get_config(self):
        return {"device": self.device, "vocab": self.vocab, "huggingface_model_id": self.huggingface_model_id,
                "pretrain_cfg": self.pretrain_cfg}

    use_cuda_graphs = True      # single-utterance calls are launch-bound: replay one CUDA graph per input length

    def _single_graphed(self, wav_input, wav_len):
        def fn(w, l):
            tv, logits, pred = self._heads(w, l)
            return tv, logits, pred, ops.softmax_rows(logits.contiguous())

        if not self.use_cuda_graphs or self.wav2vec2.precision not in ("bf16", "fp16"):
            return fn(wav_input, wav_len)
        cache = getattr(self, "_graph_cache", None)
        if cache is None:
            cache = GraphCache()
            object.__setattr__(self, "_graph_cache", cache)
        P = self.wav2vec2.plan()
        return cache.run((id(P), P.generation, int(wav_input.shape[1])), fn, wav_input, wav_len)

    def get_aptai_output(self, wav):
        """models/aptai.py:125-179 (including the (46,T,1) shape of `phn_fc_probs`, Appendix B)."""
        self.eval()
        dev = next(self.wav2vec2.parameters()).device
        with torch.no_grad():
            if type(wav) is torch.Tensor:
                wav = wav[0]
            wav_input = torch.as_tensor(np.asarray(wav), dtype=torch.float32).reshape(1, -1).to(dev)
            wav_len = torch.tensor([wav_input.shape[1]], dtype=torch.int64, device=dev)
            tv, logits, pred, probs = self._single_graphed(wav_input, wav_len)
            tvs = tv[0].cpu().numpy()
            tvs_pred = {n: list(tvs[:, i]) for i, n in enumerate(TV_NAMES)}
            return {
                "phn_fc_probs": probs.permute(2, 1, 0).cpu().numpy(),       # `.T` of a (1,T,46) tensor
                "phn_fc_logits": logits[0].cpu().numpy(),
                "phn_fc_pred": pred[0].cpu().numpy(),
                "tvs_pred": tvs_pred,
            }
