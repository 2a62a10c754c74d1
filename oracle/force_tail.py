"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the trainable tail of Force_APTAI for the
training-step parity tests — models/force_aptai.py:58-75 (modules), :116-150 (forward from the recogniser's hidden
states to the loss), models/modules.py:65-117 (ForwardSumLoss), :129-153 (CrossAttention), :190-214 (RNN, with the
intended `packed_output` semantics for batch > 1: the reference's line 207 raises NameError), :217-235 (PE).

Plain torch modules with the reference's parameter names, so `load_state_dict` takes the drop-in's tail weights and
autograd yields gradients keyed by the same names.  `reg` replays the three dropouts of a stochastic step with
explicit multiplicative masks (already scaled by 1/(1-p)): 'frame' (force_aptai.py:123), 'pe' (modules.py:235),
'rnn' (modules.py:199)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from . import heads as oh


class _XAtt(nn.Module):
    def __init__(self):
        super().__init__()
        self.q, self.k, self.layer_norm = nn.Linear(128, 128), nn.Linear(128, 128), nn.LayerNorm(256)


class _RNN(nn.Module):
    def __init__(self):
        super().__init__()
        self.lstm = nn.LSTM(256, 256, bidirectional=True, num_layers=1, batch_first=True)
        self.linear = nn.Sequential(nn.Linear(512, 256), nn.Dropout(0.0), nn.Tanh(), nn.Linear(256, 9))


class ForceTail(nn.Module):
    def __init__(self, hidden_size: int, vocab: int):
        super().__init__()
        self.xatt = _XAtt()
        self.frame_lin = nn.Linear(hidden_size, 128)
        self.phn_emb_layer = nn.Embedding(vocab, 128, padding_idx=0)
        self.rnn = _RNN()

    def forward(self, h, ids, frame_lens, phn_lens, tv_targets, reg=None):
        """h fp32 [B,T,H] (recogniser output), ids int [B,60] (0 = padding) -> dict(loss, tv_loss, align_loss, tvs)."""
        reg = reg or {}
        m = lambda key, x: x if reg.get(key) is None else x * reg[key].view(x.shape)
        B, T, _ = h.shape
        phn = self.phn_emb_layer(ids.long())                                         # force_aptai.py:118
        phn = m("pe", phn + oh.positional_encoding()[:, 0, :][None])                 # :119 (PE over the slot axis)
        fh = m("frame", self.frame_lin(h))                                           # :122-123
        mask = (ids != 0).int()
        x = self.xatt
        att_out, energy = oh.cross_attention(fh, phn, mask, x.q.weight, x.q.bias, x.k.weight, x.k.bias,
                                             x.layer_norm.weight, x.layer_norm.bias)  # :126
        att = torch.log_softmax(energy + ((1 - mask) * -1000.0).unsqueeze(1), dim=-1)  # :128-130
        if B > 1:                                                                    # modules.py:203-208
            packed = pack_padded_sequence(att_out, torch.as_tensor(frame_lens), batch_first=True, enforce_sorted=False)
            hidden, _ = pad_packed_sequence(self.rnn.lstm(packed)[0], batch_first=True, total_length=T)
        else:
            hidden, _ = self.rnn.lstm(att_out)
        lin = self.rnn.linear
        raw = lin[3](torch.tanh(m("rnn", lin[0](hidden))))                           # modules.py:198-201
        taps = oh.lowpass_taps()
        tvs = _lowpass_autograd(raw, taps)                                           # force_aptai.py:134
        tv_mask = tv_targets != -100.0
        tv_loss = F.mse_loss(tvs[tv_mask], tv_targets[tv_mask], reduction="mean")    # :137-141
        align = forward_sum_loss(att.unsqueeze(1), phn_lens, frame_lens)             # :142
        return {"loss": 0.4 * tv_loss + 0.6 * align, "tv_loss": tv_loss, "align_loss": align, "tvs": tvs, "att": att}


def _lowpass_autograd(y, taps):
    """models/modules.py:46-61 kept differentiable: per-channel conv1d('same') in float64, cast back to float32."""
    B, L, C = y.shape
    yd = y.double().permute(0, 2, 1).reshape(B * C, 1, L)
    out = F.conv1d(yd, taps.view(1, 1, -1).double(), padding="same")
    return out.view(B, C, L).permute(0, 2, 1).float()


def forward_sum_loss(attn_logprob, text_lens, mel_lens, blank_logprob=-1.0):
    """models/modules.py:79-117, torch autograd version (oracle/ctc.py holds the NumPy value-only restatement)."""
    pd = F.pad(attn_logprob, (1, 0, 0, 0, 0, 0, 0, 0), value=blank_logprob)
    ctc = nn.CTCLoss(zero_infinity=True)
    total = 0.0
    for b in range(attn_logprob.shape[0]):
        tl, ml = int(text_lens[b]), int(mel_lens[b])
        target = torch.arange(1, tl + 1).unsqueeze(0)
        cur = pd[b].permute(1, 0, 2)[:ml, :, : tl + 1]
        cur = torch.log_softmax(cur[None], dim=3)[0]
        total = total + ctc(cur, target, input_lengths=torch.tensor([ml]), target_lengths=torch.tensor([tl]))
    return total / attn_logprob.shape[0]
