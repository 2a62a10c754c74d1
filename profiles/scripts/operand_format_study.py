"""CPU study behind the fp16 operand mode: the oracle's 24x1024 forward with every GEMM / attention operand rounded to
bf16 or to IEEE fp16 (fp32 accumulate, fp32 residual stream, fp32 LayerNorm statistics - the roundings the sm_100a path
performs), compared with the unrounded fp32 forward on the phoneme argmax and the logits.  Test infrastructure only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from aptai_b200.config import W2V2Config
from aptai_b200.synth import backbone_state_dict, linear_params, waveforms
from oracle import w2v2 as ow

torch.set_num_threads(os.cpu_count())


def run(dt, sd, cfg, wav, lens):
    r = (lambda t: t) if dt is None else (lambda t: t.to(dt).float())
    lin0, mm0 = F.linear, torch.matmul

    class Shim:
        def __getattr__(self, k):
            return getattr(F, k)

        @staticmethod
        def linear(x, w, b=None):
            return lin0(r(x), r(w), b)

    def attention(sd, p, cfg, x, key_mask, prob_mask=None):
        B, T, H = x.shape
        nh = cfg.num_attention_heads
        d = H // nh
        x = r(x)
        q = r(lin0(x, r(sd[p + "q_proj.weight"]), sd[p + "q_proj.bias"])).view(B, T, nh, d).transpose(1, 2)
        k = r(lin0(x, r(sd[p + "k_proj.weight"]), sd[p + "k_proj.bias"])).view(B, T, nh, d).transpose(1, 2)
        v = r(lin0(x, r(sd[p + "v_proj.weight"]), sd[p + "v_proj.bias"])).view(B, T, nh, d).transpose(1, 2)
        s = mm0(q, k.transpose(-1, -2)) * (d ** -0.5)
        if key_mask is not None:
            s = s.masked_fill(~key_mask[:, None, None, :], float("-inf"))
        mx = s.amax(-1, keepdim=True)
        e = torch.exp(s - mx)
        o = mm0(r(e), v) / e.sum(-1, keepdim=True)          # unnormalised P rounded, fp32 row sums (flash form)
        o = r(o.transpose(1, 2).reshape(B, T, H))
        return lin0(o, r(sd[p + "out_proj.weight"]), sd[p + "out_proj.bias"])

    saveF, saveA = ow.F, ow.attention
    ow.F, ow.attention = Shim(), attention
    try:
        with torch.no_grad():
            return ow.forward(sd, cfg, wav, lens)[-1]
    finally:
        ow.F, ow.attention = saveF, saveA


def main():
    n, secs = int(sys.argv[1]) if len(sys.argv) > 1 else 4, float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
    cfg = W2V2Config.large(hidden_dropout=0.0, activation_dropout=0.0, attention_dropout=0.0, final_dropout=0.0,
                           layerdrop=0.0, apply_spec_augment=False)
    sd = backbone_state_dict(cfg, 0)
    L = int(16000 * secs)
    lens = [L - 1600 * i for i in range(n)]
    wav = waveforms(n, L, lens, seed=77)
    pw, pb = linear_params(102, 46, 1024)
    tw, tb = linear_params(101, 9, 1024)
    ref = run(None, sd, cfg, wav, lens)
    T = ref.shape[1]
    fl = torch.tensor([ow.conv_out_length(x, cfg) for x in lens])
    valid = torch.arange(T)[None] < fl[:, None]
    lref = F.linear(ref, pw, pb)
    tvref = F.linear(ref, tw, tb)
    top2 = lref.topk(2, -1).values
    print(f"frames {int(valid.sum())}, median top-1/top-2 margin {float((top2[..., 0] - top2[..., 1])[valid].median()):.3f}")
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        h = run(dt, sd, cfg, wav, lens)
        lg = F.linear(h, pw, pb)
        agree = float((lg.argmax(-1) == lref.argmax(-1))[valid].float().mean())
        print(f"{name}: argmax agreement {100 * agree:.3f} %  logits max-abs {float((lg - lref)[valid].abs().max()):.2e} "
              f"rms {float((lg - lref)[valid].pow(2).mean().sqrt()):.2e}  tv-head max-abs "
              f"{float((F.linear(h, tw, tb) - tvref)[valid].abs().max()):.2e}  |h| max {float(h.abs().max()):.1f}")


if __name__ == "__main__":
    main()
