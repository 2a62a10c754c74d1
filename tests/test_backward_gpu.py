"""Parity of the training-step kernels (dgrad epilogues, tcgen05 wgrad, attention backward, LayerNorm / heads / loss
backward, weight-norm backward, fused Adam) against torch fp32 autograd on the same (bf16-rounded) operands.
Calls go through the C ABI (aptai_b200.ops -> ctypes)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from aptai_b200 import ops


def _rand(shape, dev, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dev)


def _rel(out, ref):
    return ((out.float() - ref.float()).norm() / ref.float().norm().clamp_min(1e-30)).item()


def test_linear_out_pre_and_gelu_dgrad(cuda):
    M, K, N = 700, 1024, 4096
    a = _rand((M, K), cuda, 1.0, 1).bfloat16()
    w = _rand((N, K), cuda, 0.03, 2).bfloat16()
    b = _rand((N,), cuda, 0.3, 3)
    pre = torch.empty((M, N), dtype=torch.bfloat16, device=cuda)
    _, g = ops.linear(a, w, b, act=1, out_pre=pre)
    ref_pre = a.float() @ w.float().t() + b
    assert _rel(pre, ref_pre) < 5e-3
    assert _rel(g, F.gelu(ref_pre)) < 5e-3
    # dgrad through the GELU: dU = (dG @ W2) * gelu'(u)
    dy = _rand((M, K), cuda, 1.0, 4).bfloat16()            # gradient of a [M, K] output of a second Linear(N -> K)
    w2t = _rand((N, K), cuda, 0.03, 5).bfloat16()           # = W2^T, [N, K]: dG = dy @ W2 = dy @ w2t^T
    _, du = ops.linear(dy, w2t, None, act=2, aux=pre)
    u = pre.float().requires_grad_(True)
    F.gelu(u).backward(dy.float() @ w2t.float().t())
    assert _rel(du, u.grad) < 6e-3


@pytest.mark.parametrize("M,N,K", [(12768, 1024, 4096), (12768, 4096, 1024), (5000, 3072, 1024), (777, 768, 768),
                                   (3990, 1024, 512), (64, 128, 256), (100, 2304, 768)])
def test_wgrad(cuda, M, N, K):
    dy = _rand((M, N), cuda, 1.0, 1).bfloat16()
    x = _rand((M, K), cuda, 1.0, 2).bfloat16()
    dw = torch.zeros((N, K), dtype=torch.float32, device=cuda)
    ops.wgrad(dy, x, dw)
    ref = dy.float().t() @ x.float()
    assert _rel(dw, ref) < 2e-5, _rel(dw, ref)
    ops.wgrad(dy, x, dw)                      # accumulates
    assert _rel(dw, 2 * ref) < 2e-5


@pytest.mark.parametrize("B,T,H,groups", [(3, 210, 1024, 16), (2, 99, 768, 16)])
def test_posconv_backward(cuda, B, T, H, groups):
    taps, gw = 128, H // groups
    x = _rand((B, T, H), cuda, 1.0, 1)
    v = _rand((H, gw, taps), cuda, 0.05, 2)
    g = v.norm(dim=(0, 1), keepdim=True) * 1.1
    bias = _rand((H,), cuda, 0.1, 3)
    dy = _rand((B, T, H), cuda, 1.0, 4).bfloat16()
    # reference: weight-normed grouped conv on bf16-rounded operands, fp32 math
    xb = x.bfloat16().float().requires_grad_(True)
    gr, vr = g.clone().requires_grad_(True), v.clone().requires_grad_(True)
    w = gr * vr / vr.norm(dim=(0, 1), keepdim=True)
    wq = (w.bfloat16().float() - w).detach() + w          # bf16-rounded values, identity gradient
    y = F.conv1d(xb.transpose(1, 2), wq, None, padding=taps // 2, groups=groups)[:, :, :T].transpose(1, 2)
    y.backward(dy.float())
    # ours: forward-layout operands
    xp = ops.cast_pad(x, taps // 2)
    dwf = ops.posconv_wgrad(dy, xp, groups, taps)
    dg = torch.zeros_like(g)
    dv = torch.zeros_like(v)
    ops.posconv_weightnorm_bwd(dwf, g.contiguous(), v.contiguous(), dg, dv)
    assert _rel(dv, vr.grad) < 1e-3, _rel(dv, vr.grad)
    assert _rel(dg, gr.grad) < 1e-3, _rel(dg, gr.grad)
    # dgrad = the forward kernel on the flipped, in/out-swapped weight, reading dy one row later
    vt = v.view(groups, gw, gw, taps).permute(0, 2, 1, 3).flip(-1).reshape(H, gw, taps).contiguous()
    wt = ops.posconv_fold(g.flip(-1).contiguous(), vt, cpad=64)
    dyp = ops.cast_pad(dy.float(), taps // 2)
    dx = torch.zeros((B * T, H), dtype=torch.float32, device=cuda)
    ops.posconv(dyp, wt, None, None, T, H, groups, taps, dx, act=0, row_shift=1)
    assert _rel(dx.view(B, T, H), xb.grad) < 6e-3, _rel(dx.view(B, T, H), xb.grad)


@pytest.mark.parametrize("B,T,heads,lens", [(2, 399, 16, [399, 250]), (3, 130, 12, [130, 1, 77]),
                                             (1, 999, 16, [999]), (4, 64, 16, [64, 64, 30, 5])])
def test_attention_backward(cuda, B, T, heads, lens):
    H = heads * 64
    M = B * T
    qkv = _rand((M, 3 * H), cuda, 1.0, 1)
    qkv[:, :H] *= 0.125 * 2.0           # the forward stores q pre-scaled
    qkv = qkv.bfloat16()
    klen = torch.tensor(lens, dtype=torch.int32, device=cuda)
    lse = torch.empty((B, heads, T), dtype=torch.float32, device=cuda)
    ctx = ops.attention(qkv, klen, B, T, heads, lse=lse)
    d_ctx = _rand((M, H), cuda, 1.0, 2).bfloat16()
    dqkv = ops.attention_bwd(qkv, ctx, d_ctx, lse, klen, B, T, heads, q_scale=1.0)
    # reference
    x = qkv.float().view(B, T, 3, heads, 64).requires_grad_(True)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    mask = torch.arange(T, device=cuda)[None, :] >= klen[:, None].clamp_min(1)
    s = s.masked_fill(mask[:, None, None, :], float("-inf"))
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(M, H)
    assert _rel(ctx, o) < 1e-2
    ref_lse = torch.logsumexp(s, -1) * math.log2(math.e)
    torch.testing.assert_close(lse, ref_lse, atol=2e-2, rtol=1e-3)
    o.backward(d_ctx.float())
    ref = x.grad.reshape(M, 3 * H)
    for name, sl in (("dq", slice(0, H)), ("dk", slice(H, 2 * H)), ("dv", slice(2 * H, 3 * H))):
        r = _rel(dqkv[:, sl], ref[:, sl])
        assert r < 1.5e-2, (name, r)
    # padded keys receive exactly zero gradient
    for b, n in enumerate(lens):
        assert dqkv[b * T + n:(b + 1) * T, H:].abs().max().item() == 0 if n < T else True


@pytest.mark.parametrize("rows,cols", [(1000, 1024), (333, 768), (4100, 512)])
def test_layernorm_bwd(cuda, rows, cols):
    x = _rand((rows, cols), cuda, 2.0, 1) + 0.5
    dy = _rand((rows, cols), cuda, 1.0, 2)
    gamma = _rand((cols,), cuda, 0.3, 3) + 1.0
    beta = _rand((cols,), cuda, 0.3, 4)
    dres = _rand((rows, cols), cuda, 1.0, 5)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (cols,), gr, br, 1e-5).backward(dy)
    dg = torch.zeros_like(gamma)
    db = torch.zeros_like(beta)
    dx, dxb = ops.layernorm_bwd(dy, x, gamma, 1e-5, dres=dres, dgamma=dg, dbeta=db, want_bf16=True)
    torch.testing.assert_close(dx, xr.grad + dres, atol=2e-5, rtol=1e-4)
    assert _rel(dxb, xr.grad + dres) < 4e-3
    torch.testing.assert_close(dg, gr.grad, atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(db, br.grad, atol=2e-3, rtol=1e-4)
    # fused column sums of the output gradient (the bias gradient of the Linear that wrote x): accumulate into dcolsum,
    # same dx / dgamma / dbeta as without it
    dc = torch.full((cols,), 0.25, device=cuda)
    dg2, db2 = torch.zeros_like(gamma), torch.zeros_like(beta)
    dx2, dxb2 = ops.layernorm_bwd(dy, x, gamma, 1e-5, dres=dres, dgamma=dg2, dbeta=db2, want_bf16=True, dcolsum=dc)
    assert torch.equal(dx2, dx) and torch.equal(dxb2, dxb)
    torch.testing.assert_close(dg2, dg, atol=2e-3, rtol=1e-4)
    torch.testing.assert_close(dc, 0.25 + (xr.grad + dres).sum(0), atol=3e-3, rtol=1e-4)
    dc2 = torch.zeros((cols,), device=cuda)
    ops.layernorm_bwd(dy, x, gamma, 1e-5, want_bf16=False, dcolsum=dc2)          # no residual, no dgamma / dbeta
    torch.testing.assert_close(dc2, xr.grad.sum(0), atol=3e-3, rtol=1e-4)


def test_colsum(cuda):
    x = _rand((12768, 4096), cuda, 1.0, 1).bfloat16()
    out = torch.zeros((4096,), dtype=torch.float32, device=cuda)
    ops.colsum(x, out)
    torch.testing.assert_close(out, x.float().sum(0), atol=2e-2, rtol=1e-4)
    y = _rand((700, 46 * 2), cuda, 1.0, 2)
    out = torch.ones((92,), dtype=torch.float32, device=cuda)
    ops.colsum(y, out, scale=0.5)
    torch.testing.assert_close(out, 1 + 0.5 * y.sum(0), atol=1e-3, rtol=1e-4)


def test_heads_and_loss_bwd(cuda):
    rows, H, B, T = 2 * 150, 1024, 2, 150
    h = _rand((rows, H), cuda, 1.0, 1)
    wa, ba = _rand((9, H), cuda, 0.03, 2), _rand((9,), cuda, 0.1, 3)
    wb, bb = _rand((46, H), cuda, 0.03, 4), _rand((46,), cuda, 0.1, 5)
    taps = torch.hann_window(51, periodic=False, dtype=torch.float64, device=cuda)
    taps = (taps / taps.sum()).contiguous()
    g = torch.Generator().manual_seed(6)
    phn = torch.randint(1, 46, (B, T), generator=g).to(cuda)
    phn[1, 100:] = 0
    tgt = _rand((B, T, 9), cuda, 1.0, 7)
    tgt[1, 100:] = -100.0
    # reference (models/aptai.py:83-102)
    hr = h.clone().requires_grad_(True)
    war, bar, wbr, bbr = (t.clone().requires_grad_(True) for t in (wa, ba, wb, bb))
    tv = torch.tanh(hr) @ war.t() + bar
    tv_lp = F.conv1d(tv.view(B, T, 9).transpose(1, 2).double().reshape(B * 9, 1, T), taps.view(1, 1, -1),
                     padding="same").float().view(B, 9, T).transpose(1, 2)
    logits = F.leaky_relu(hr) @ wbr.t() + bbr
    m = tgt != -100.0
    mse = F.mse_loss(tv_lp[m], tgt[m])
    pm = (phn != 0).flatten()
    ce = F.cross_entropy(logits[pm], phn.flatten()[pm], ignore_index=0)
    loss = 0.5 * mse + 0.5 * ce
    loss.backward()
    # ours
    tv_o, lg_o, _ = ops.heads(h, wa, ba, ops.ACT_TANH, wb, bb, ops.ACT_LEAKY)
    tv_lp_o = ops.lowpass(tv_o.view(B, T, 9), taps)
    out3, ws = ops.masked_mse_ce(tv_lp_o.view(rows, 9), tgt.view(rows, 9).contiguous(), lg_o, phn.flatten(),
                                 return_ws=True)
    torch.testing.assert_close(out3[0], loss.detach(), atol=1e-5, rtol=1e-5)
    d_tvlp, d_lg = ops.masked_mse_ce_bwd(tv_lp_o.view(rows, 9), tgt.view(rows, 9).contiguous(), lg_o, phn.flatten(), ws)
    d_tv = ops.lowpass(d_tvlp.view(B, T, 9), taps).view(rows, 9)      # symmetric taps: the adjoint is the filter
    dwa, dba, dwb, dbb = (torch.zeros_like(t) for t in (wa, ba, wb, bb))
    dh = ops.heads_bwd(h, d_tv, wa, ops.ACT_TANH, dwa, dba, d_lg, wb, ops.ACT_LEAKY, dwb, dbb)
    torch.testing.assert_close(dh, hr.grad, atol=1e-7, rtol=2e-4)
    torch.testing.assert_close(dwa, war.grad, atol=1e-6, rtol=2e-4)
    torch.testing.assert_close(dba, bar.grad, atol=1e-6, rtol=2e-4)
    torch.testing.assert_close(dwb, wbr.grad, atol=1e-6, rtol=2e-4)
    torch.testing.assert_close(dbb, bbr.grad, atol=1e-6, rtol=2e-4)


def test_fused_adam_matches_torch(cuda):
    from aptai_b200.train import FusedAdam, GradBuffer
    shapes = [(1024, 1024), (1024,), (46, 1024), (3,), (70000,)]
    ps = [torch.nn.Parameter(_rand(s, cuda, 1.0, i)) for i, s in enumerate(shapes)]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    GradBuffer([(f"p{i}", p) for i, p in enumerate(ps)])
    opt = FusedAdam(ps, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    ropt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    for step in range(3):
        for i, (p, r) in enumerate(zip(ps, ref)):
            gr = _rand(p.shape, cuda, 0.1, 100 * step + i)
            p.grad.copy_(gr)
            r.grad = gr.clone()
        opt.step()
        ropt.step()
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.detach(), r.detach(), atol=1e-6, rtol=1e-5)


def test_fused_adam_consumes_pending_regions(cuda):
    """Optimizer overlap on ONE GPU: all-reduce regions registered as pending on the gradient buffer (here with
    recording no-op waits, in last-region-first order and with one tensor left uncovered) are updated region by region
    — every parameter exactly once, each region only after ITS wait — and the result equals torch.optim.Adam."""
    from aptai_b200.train import FusedAdam, GradBuffer
    shapes = [(1024, 1024), (1024,), (46, 1024), (3,), (200000,), (64, 64)]
    ps = [torch.nn.Parameter(_rand(s, cuda, 1.0, i)) for i, s in enumerate(shapes)]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    gb = GradBuffer([(f"p{i}", p) for i, p in enumerate(ps)])
    opt = FusedAdam(ps, lr=1e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.0)
    ropt = torch.optim.Adam(ref, lr=1e-3, betas=(0.9, 0.98), eps=1e-8)
    o = gb.offsets
    regions = [(o["p4"], o["p5"]), (o["p2"], o["p4"]), (0, o["p2"])]        # p5 is covered by no region
    for step in range(3):
        for i, (p, r) in enumerate(zip(ps, ref)):
            gr = _rand(p.shape, cuda, 0.1, 100 * step + i)
            p.grad.copy_(gr)
            r.grad = gr.clone()
        calls = []
        before = [p.detach().clone() for p in ps]
        gb.pending = [((lambda k=k: calls.append(k)), lo, hi) for k, (lo, hi) in enumerate(regions)]
        opt.step()
        ropt.step()
        assert calls == [0, 1, 2] and gb.pending == []
        assert all(not torch.equal(b, p.detach()) for b, p in zip(before, ps))     # every tensor moved (p5 too)
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.detach(), r.detach(), atol=1e-6, rtol=1e-5)
    # wait_pending() / zero_grad() complete what is left without an optimizer step
    calls = []
    gb.pending = [((lambda: calls.append("w")), 0, gb.numel)]
    opt.zero_grad()
    assert calls == ["w"] and gb.pending == [] and float(gb.flat.abs().sum()) == 0.0


@pytest.mark.parametrize("half", [0, 1])
def test_prepare_weights_table(cuda, half):
    """One-launch operand preparation (aptai_prepare_weights_fmt): 16-bit [N][K] copies with a scale and a row offset
    inside a fused buffer, transposed copies with their own pitch / offset / scale, fused fp32 biases — vector path
    (shapes and pitches multiples of 4) and scalar path (odd shapes), bf16 and fp16, bit-exact against torch."""
    from aptai_b200.backbone import _WeightTable
    dt = torch.float16 if half else torch.bfloat16
    H = 192
    wq, wk = _rand((H, H), cuda, 1.0, 1), _rand((H, H), cuda, 1.0, 2)
    bq, bk = _rand((H,), cuda, 1.0, 3), _rand((H,), cuda, 1.0, 4)
    w1 = _rand((320, 100), cuda, 1.0, 5)              # 100 columns: a partial 64-wide tile on the vector path
    wodd = _rand((37, 53), cuda, 1.0, 6)              # scalar path
    qk = torch.full((2 * H, H), 7.0, dtype=dt, device=cuda)
    qkt = torch.full((H, 2 * H), 7.0, dtype=dt, device=cuda)
    qkb = torch.full((2 * H,), 7.0, device=cuda)
    w1c, w1t = torch.empty((320, 100), dtype=dt, device=cuda), torch.empty((100, 320), dtype=dt, device=cuda)
    oddc, oddt = torch.empty((37, 53), dtype=dt, device=cuda), torch.empty((53, 37), dtype=dt, device=cuda)
    tb = _WeightTable(cuda)
    tb.add(wq, dst=qk, dst_off=0, dst_ld=H, scale=0.125, dst_t=qkt, dst_t_off=0, dst_t_ld=2 * H, scale_t=1.0)
    tb.add(bq, dst_f32=qkb, f32_off=0, scale=0.125)
    tb.add(wk, dst=qk, dst_off=H * H, dst_ld=H, dst_t=qkt, dst_t_off=H, dst_t_ld=2 * H)
    tb.add(bk, dst_f32=qkb, f32_off=H)
    tb.add(w1, dst=w1c, dst_t=w1t)
    tb.add(wodd, dst=oddc, dst_t=oddt, scale=0.5, scale_t=2.0)
    tb.run(half)
    torch.cuda.synchronize()
    assert torch.equal(qk, torch.cat([(wq * 0.125).to(dt), wk.to(dt)]))
    assert torch.equal(qkt, torch.cat([wq.t().to(dt), wk.t().to(dt)], dim=1))
    assert torch.equal(qkb, torch.cat([bq * 0.125, bk]))
    assert torch.equal(w1c, w1.to(dt)) and torch.equal(w1t, w1.t().to(dt))
    assert torch.equal(oddc, (wodd * 0.5).to(dt)) and torch.equal(oddt, (wodd.t() * 2.0).to(dt))
