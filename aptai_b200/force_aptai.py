"""Drop-in for the reference's models/force_aptai.py (class Force_APTAI) on the aptai_b200 kernels.

Frozen Wav2Vec2_PR -> CTC-decoded phoneme sequence (padded to 60) -> embedding + sinusoidal PE -> cross-attention
frames x phonemes -> log-softmax alignment matrix -> BiLSTM -> 9 TVs -> FIR low-pass; loss = 0.4 MSE + 0.6
forward-sum (CTC over the attention matrix).  Additive API: `forced_align()` — the CTC-Viterbi alignment of the
north star, which the reference does not have — and `phn_seqs=` to inject known phoneme
sequences instead of decoding (BASELINE config 3).
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .aptai import TV_NAMES
from .modules import CrossAttention, ForwardSumLoss, LowPassFilterLayer, PositionalEncoding, RNN
from .w2v2_pr import Wav2Vec2_PR


class Force_APTAI(nn.Module):
    def __init__(self, pr_model_path, device, vocab, w2v2_pr: Wav2Vec2_PR = None):
        super().__init__()
        self.vocab = vocab
        self.device = device
        self.hidden_drop = 0.2
        self.rnn_drop = 0.1
        self.max_phn_seq_len = 60
        self.frame_hidden_dim = 128
        self.phn_hidden_dim = 128
        self.att_hidden_dim = 128
        self.rnn_in_dim = 2 * self.att_hidden_dim
        self.pr_model_path = pr_model_path
        if w2v2_pr is None:
            assert os.path.exists(pr_model_path)
            pr_ckpt_path = os.path.join(pr_model_path, "best-model-ckpt")
            self.w2v2_pr_cfg = pickle.load(open(os.path.join(pr_ckpt_path, "model_cfg.pkl"), "rb"))
            self.w2v2_pr = Wav2Vec2_PR(self.w2v2_pr_cfg["pretrain_cfg"], self.w2v2_pr_cfg["cache_dir"],
                                       self.w2v2_pr_cfg["huggingface_model_id"], vocab).to(self.device)
            self.w2v2_pr.load_state_dict(torch.load(os.path.join(pr_ckpt_path, "pytorch_model.bin"),
                                                    map_location=torch.device(self.device)))
        else:
            self.w2v2_pr_cfg = w2v2_pr.get_config()
            self.w2v2_pr = w2v2_pr.to(self.device)
        H = self.w2v2_pr.wav2vec2.cfg.hidden_size
        self.xatt = CrossAttention(self.frame_hidden_dim, self.phn_hidden_dim, self.att_hidden_dim)
        self.align_loss = ForwardSumLoss()
        self.frame_lin = nn.Linear(H, self.frame_hidden_dim)
        self.frame_drop = nn.Dropout(self.hidden_drop)
        self.phn_emb_layer = nn.Embedding(len(self.vocab), self.phn_hidden_dim, padding_idx=0)
        self.pe_phn = PositionalEncoding(self.phn_hidden_dim, max_len=60, dropout=self.hidden_drop)
        self.rnn = RNN(self.rnn_in_dim, 9, self.rnn_drop)
        self.tv_lowpass = LowPassFilterLayer(self.device, 10, 49, 9)
        for param in self.w2v2_pr.parameters():
            param.requires_grad = False

    # ---------------------------------------------------------------------------------------------- shared trunk
    @torch.no_grad()
    def _trunk(self, audio_inputs, audio_lengths, phn_seqs=None):
        if self.training:
            raise NotImplementedError("aptai_b200: training-mode dropout/backward is not built yet; call .eval()")
        pr = self.w2v2_pr
        _, h, logits = pr._logits(audio_inputs, audio_lengths)
        dev = h.device
        frame_seq_lens = pr.wav2vec2._get_feat_extract_output_lengths(audio_lengths.reshape(-1)).cpu().tolist()
        phn_pred_list = [np.asarray(s) for s in phn_seqs] if phn_seqs is not None else pr._decode(logits)
        phn_seq_lens = [len(l) for l in phn_pred_list]
        padded = []
        for lst in phn_pred_list:
            assert len(lst) < self.max_phn_seq_len, "Need longer max phoneme sequence length."
            padded.append(np.pad(lst, (0, self.max_phn_seq_len - len(lst)), mode="constant"))
        phn_pred_seq = torch.tensor(np.stack(padded), dtype=torch.int32, device=dev)
        phn_pred_mask = (phn_pred_seq != 0).to(torch.int)
        B, T, H = h.shape
        fl = self.frame_lin                 # Linear(H,128) on the tcgen05 GEMM, bf16x3 split (fp32-accurate)
        fh = ops.linear_f32x3(h.reshape(B * T, H).contiguous(), fl.weight, fl.bias.detach().float().contiguous())
        # fused cross-attention kernel: embedding + PE, q/k projections, masked softmax, LayerNorm(256) and the
        # log-softmax alignment matrix (eval mode: the two dropouts are identities)
        xa = self.xatt
        att_out, energy, att = ops.cross_attention(fh.view(B, T, -1), phn_pred_seq.contiguous(),
                                                   self.phn_emb_layer.weight, self.pe_phn.pe[:, 0, :], xa.q.weight,
                                                   xa.q.bias, xa.k.weight, xa.k.bias, xa.layer_norm.weight,
                                                   xa.layer_norm.bias, xa.layer_norm.eps)
        return dict(h=h, logits=logits, frame_seq_lens=frame_seq_lens, phn_pred_list=phn_pred_list,
                    phn_seq_lens=phn_seq_lens, phn_pred_seq=phn_pred_seq, att_out=att_out, att=att)

    def forward(self, epoch, audio_inputs, audio_lengths, phoneme_labels, phn_frames_49hz, LA, LP, JA, TTCL, TTCD,
                TMCL, TMCD, TBCL, TBCD, phn_seqs=None):
        """models/force_aptai.py:80-178."""
        tv_targets = torch.stack([LA, LP, JA, TTCL, TTCD, TMCL, TMCD, TBCL, TBCD], dim=-1).float()
        t = self._trunk(audio_inputs, audio_lengths, phn_seqs)
        att = t["att"]
        dev = att.device
        rnn_out = self.rnn(t["att_out"], t["frame_seq_lens"])
        tvs_out = self.tv_lowpass(rnn_out[0].contiguous())
        tv_targets = tv_targets.to(dev)
        tv_pad_mask = tv_targets != -100.0
        d = (tvs_out - tv_targets)[tv_pad_mask]
        tv_loss = (d * d).mean()
        align_loss = self.align_loss(att.unsqueeze(1), t["phn_seq_lens"], t["frame_seq_lens"])
        a = 0.4
        loss = a * tv_loss + (1 - a) * align_loss
        align_out = torch.max(att, axis=2)[1]
        frame_phn = torch.gather(t["phn_pred_seq"].long(), 1, align_out).cpu().numpy()      # one D->H copy
        pred_frame_phns = [[int(x) for x in frame_phn[b, : t["frame_seq_lens"][b]]] for b in range(att.shape[0])]
        return {"loss": loss, "tv_loss": tv_loss, "align_loss": align_loss, "tvs_pred": tvs_out,
                "pred_frame_phns": pred_frame_phns, "pred_ctc_phn_seq": t["phn_pred_list"]}

    def get_config(self):
        return {"pr_model_path": self.pr_model_path, "w2v2_pr_cfg": self.w2v2_pr_cfg, "device": self.device,
                "vocab": self.vocab}

    def _single(self, wav):
        dev = next(self.w2v2_pr.parameters()).device
        if type(wav) is torch.Tensor:
            wav = wav[0]
        wav_input = torch.as_tensor(np.asarray(wav), dtype=torch.float32).reshape(1, -1).to(dev)
        wav_len = torch.tensor([wav_input.shape[1]], dtype=torch.int64, device=dev)
        return wav_input, wav_len

    def get_alignment(self, wav, phn_seq=None):
        """models/force_aptai.py:188-236: (N, T) log-softmax alignment matrix."""
        self.eval()
        wav_input, wav_len = self._single(wav)
        t = self._trunk(wav_input, wav_len, None if phn_seq is None else [phn_seq])
        att = t["att"][0]
        res = att[0: t["frame_seq_lens"][0], 0: t["phn_seq_lens"][0]].permute(1, 0)
        return {"alignment": res.detach().cpu().numpy()}

    def get_faptai_output(self, wav, phn_seq=None):
        """models/force_aptai.py:238-322."""
        self.eval()
        wav_input, wav_len = self._single(wav)
        t = self._trunk(wav_input, wav_len, None if phn_seq is None else [phn_seq])
        with torch.no_grad():
            rnn_out = self.rnn(t["att_out"], t["frame_seq_lens"])
            tvs = self.tv_lowpass(rnn_out[0].contiguous())[0].cpu().numpy()
        tvs_pred = {n: list(tvs[:, i]) for i, n in enumerate(TV_NAMES)}
        align_out = torch.max(t["att"][0], axis=1)[1]
        pred_frame_phns = [int(x) for x in t["phn_pred_seq"][0].long()[align_out].cpu().numpy()]
        return {"tvs_pred": tvs_pred, "pred_frame_phns": pred_frame_phns, "pred_ctc_phn_seq": t["phn_pred_list"],
                "hidden_alignment": t["att_out"], "hidden_tvs": rnn_out[1]}

    # ---------------------------------------------------------------------------------------------- additive API
    @torch.no_grad()
    def forced_align(self, audio_inputs, audio_lengths, phn_seqs, log_probs=None):
        """CTC-Viterbi forced alignment of known phoneme sequences against the frozen recogniser's log-probs
        (bit-exact with torchaudio.functional.forced_align on identical log-probs).  Returns (paths int32 [B,T]
        with -1 beyond each utterance, scores fp32 [B,T], status int32 [B])."""
        pr = self.w2v2_pr
        dev = next(pr.parameters()).device
        if log_probs is None:
            _, _, logits = pr._logits(audio_inputs, audio_lengths)
            log_probs = ops.softmax_rows(logits.contiguous(), log=True)
        B, T, V = log_probs.shape
        flen = pr.wav2vec2._get_feat_extract_output_lengths(audio_lengths.reshape(-1).to(dev)).to(torch.int32)
        Smax = max(1, max(len(s) for s in phn_seqs))
        tg = np.zeros((B, Smax), dtype=np.int32)
        for b, s in enumerate(phn_seqs):
            tg[b, : len(s)] = np.asarray(s, dtype=np.int32)
        tl = torch.tensor([len(s) for s in phn_seqs], dtype=torch.int32, device=dev)
        return ops.ctc_viterbi(log_probs.contiguous(), torch.from_numpy(tg).to(dev), flen.contiguous(), tl,
                               blank=int(pr.wav2vec2.cfg.blank))
