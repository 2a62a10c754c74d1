"""End-to-end parity of the CUDA encoder (aptai_b200.backbone) against the CPU oracle on identical weights."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from aptai_b200.backbone import Wav2Vec2Backbone
from aptai_b200.config import W2V2Config
from oracle import w2v2 as ow
from oracle.weights import backbone_state_dict, waveforms


def _cfg(variant, **kw):
    base = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512,
                num_conv_pos_embedding_groups=4)
    base.update(kw)
    if variant == "layer":
        return W2V2Config.large(**base)
    return W2V2Config.base(**base)


def _run(cfg, lens, L, cuda):
    sd = backbone_state_dict(cfg, seed=0)
    wav = waveforms(len(lens), L, lens, seed=1234)
    ref = ow.forward(sd, cfg, wav, lens)
    m = Wav2Vec2Backbone(cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    out = m(wav.to(cuda), attention_mask=torch.tensor(lens, device=cuda)[:, None], output_hidden_states=True)
    return ref, out


@pytest.mark.parametrize("variant", ["layer", "group"])
def test_tiny_encoder_matches_oracle(cuda, variant):
    cfg = _cfg(variant)
    lens = [16000, 12000, 9000]
    ref, out = _run(cfg, lens, 16000, cuda)
    assert len(out.hidden_states) == cfg.num_hidden_layers + 1
    for i, (r, o) in enumerate(zip(ref, out.hidden_states)):
        err = (o.cpu() - r).abs().max().item()
        assert err < 6e-2, f"{variant}: hidden[{i}] max abs err {err}"
    # padded frames are zeroed before the encoder but still computed afterwards (HF:679-682): compare all frames
    err = (out.last_hidden_state.cpu() - ref[-1]).abs()
    assert err.mean().item() < 1e-2     # bf16 operands: SURVEY.md Appendix D measures hidden max|d| 2.4e-2..4e-2


@pytest.mark.parametrize("variant,H,heads,F", [("layer", 1024, 16, 4096), ("group", 768, 12, 3072)])
def test_full_width_two_layers(cuda, variant, H, heads, F):
    """Real hidden sizes (pos-conv group widths 64 and 48, N tiles of 256) with two layers to keep the CPU oracle fast."""
    cfg = _cfg(variant, hidden_size=H, num_attention_heads=heads, intermediate_size=F, num_hidden_layers=2,
               num_conv_pos_embedding_groups=16)
    lens = [32000, 20000]
    ref, out = _run(cfg, lens, 32000, cuda)
    err = (out.last_hidden_state.cpu() - ref[-1]).abs()
    assert err.max().item() < 6e-2 and err.mean().item() < 1e-2, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("variant", ["layer", "group"])
def test_edge_lengths_ragged_and_minimal(cuda, variant):
    """Ragged batch whose shortest utterance yields a single frame (400 samples -> T=1), next to a full-length one."""
    cfg = _cfg(variant)
    lens = [8000, 400, 3217]
    ref, out = _run(cfg, lens, 8000, cuda)
    T = ref[-1].shape[1]
    assert T == 24 and out.last_hidden_state.shape == ref[-1].shape
    err = (out.last_hidden_state.cpu() - ref[-1]).abs()
    assert err.max().item() < 8e-2 and err.mean().item() < 1.2e-2, (err.max().item(), err.mean().item())


def test_single_frame_utterance(cuda):
    """The shortest legal input: 400 samples = one frame, batch of one (every conv layer produces 79..1 frames)."""
    cfg = _cfg("layer")
    ref, out = _run(cfg, [400], 400, cuda)
    assert out.last_hidden_state.shape == (1, 1, cfg.hidden_size)
    assert (out.last_hidden_state.cpu() - ref[-1]).abs().max().item() < 8e-2


def test_maximum_length_20s(cuda):
    """20 s = 320 000 samples -> 999 frames (8 attention KV tiles of 128 queries, 16 of 64 keys), padded partner of 2 s."""
    cfg = _cfg("layer")
    lens = [320000, 32000]
    ref, out = _run(cfg, lens, 320000, cuda)
    assert ref[-1].shape[1] == 999
    err = (out.last_hidden_state.cpu() - ref[-1]).abs()
    assert err.max().item() < 8e-2 and err.mean().item() < 1.2e-2, (err.max().item(), err.mean().item())
    # valid frames of the short utterance must not depend on the 900 padded frames behind it ('layer' variant,
    # SURVEY.md fact 7): run the same waveform alone through the same model
    sd = backbone_state_dict(cfg, seed=0)
    m = Wav2Vec2Backbone(cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    wav = waveforms(2, 320000, lens, seed=1234)
    alone = m(wav[1:2, :32000].contiguous().to(cuda), attention_mask=torch.tensor([[32000]], device=cuda))
    d = (out.last_hidden_state[1, :99] - alone.last_hidden_state[0, :99]).abs().max().item()
    assert d < 3e-2, d


@pytest.mark.parametrize("variant,H,heads,F,lens", [("layer", 1024, 16, 4096, [64000, 20000]),
                                                    ("group", 768, 12, 3072, [32000, 20000]),
                                                    ("layer", 256, 4, 512, [8000, 400, 3217])])
def test_fp16_operand_mode(cuda, variant, H, heads, F, lens):
    """precision="fp16": the same kernels on IEEE fp16 operands (both LayerNorm wirings, both pos-conv paths, both
    attention kernels: T = 199 / 99 / 24) — the hidden states land eight times closer to the fp32 oracle than in bf16,
    switching back gives the bf16 result bit for bit, and training stays on the bf16 plan."""
    cfg = _cfg(variant, hidden_size=H, num_attention_heads=heads, intermediate_size=F,
               num_conv_pos_embedding_groups=16 if H > 256 else 4)
    sd = backbone_state_dict(cfg, seed=0)
    wav = waveforms(len(lens), max(lens), lens, seed=1234)
    ref = ow.forward(sd, cfg, wav, lens)
    m = Wav2Vec2Backbone(cfg)
    m.load_state_dict(sd, strict=True)
    m = m.to(cuda).eval()
    am = torch.tensor(lens, device=cuda)[:, None]
    o_bf = m(wav.to(cuda), attention_mask=am).last_hidden_state.clone()
    m.set_precision("fp16")
    out = m(wav.to(cuda), attention_mask=am, output_hidden_states=True)
    assert m._plan_h is not None and m._plan_h.layers[0].qkv_w.dtype == torch.float16
    assert m._plan.layers[0].qkv_w.dtype == torch.bfloat16
    for i, (r, o) in enumerate(zip(ref, out.hidden_states)):
        err = (o.cpu() - r).abs().max().item()
        assert err < 8e-3, f"{variant}: hidden[{i}] max abs err {err}"
    err = (out.last_hidden_state.cpu() - ref[-1]).abs()
    e_bf = (o_bf.cpu() - ref[-1]).abs()
    assert err.mean().item() < 1.5e-3 and err.mean().item() < 0.5 * e_bf.mean().item(), (err.mean().item(),
                                                                                            e_bf.mean().item())
    m.set_precision("bf16")
    assert torch.equal(m(wav.to(cuda), attention_mask=am).last_hidden_state, o_bf)
    with pytest.raises(ValueError):
        m.set_precision("fp8")
