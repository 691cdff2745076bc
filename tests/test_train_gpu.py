"""Training step: forward in training mode + back-propagation through time on the GPU against the
reference's own arithmetic -- torch's CPU ``nn.LSTM`` under autograd (oracle/lstm_oracle.training_reference).

Tolerance: fp32 with a different summation order (the gradients are sums over B*T terms): loss within 1e-5 relative,
every gradient within 2e-4 of its tensor's max-abs against the reference (measured: ~1e-6).  "The reference" is its fp32 run
or, per tensor, its float64 run: a ReLU input within rounding of zero switches sides between two roundings of the same
forward and moves the gradient discontinuously -- on the 9 x 64 case the reference's own fp32 and float64 runs differ by 9e-3
for that reason, and the GPU forward (tcgen05 path: split-fp16 operands, approximate exp2) lands on the float64 side there.
"""
import numpy as np
import pytest
import torch

from oracle import lstm_oracle as lo

pytestmark = pytest.mark.gpu

REL_TOL = 2e-4


@pytest.fixture(scope="module")
def lib(built):
    from hss import _lib

    assert torch.cuda.is_available()
    return _lib.lib()


def make_model(seed, F, B, H):
    from hss.model.segmenter import HeartSoundSegmenter

    torch.manual_seed(seed)
    return HeartSoundSegmenter(input_size=F, batch_size=B, hidden_size=H)


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


@pytest.mark.parametrize("B,T,F,H", [(3, 17, 44, 240), (5, 40, 7, 32), (9, 64, 44, 240), (1, 1, 44, 240)])
def test_training_step_gradients_match_torch_cpu(lib, B, T, F, H):
    m = make_model(B + T, F, B, H).cuda().train()
    m.dropout.p = 0.0                                         # dropout is covered by the next test
    params, h0, c0 = lo.reference_params(B + T, F, B, H)
    g = torch.Generator().manual_seed(T)
    x = torch.randn(B, T, F, generator=g)
    y = torch.randint(0, 4, (B, T), generator=g)
    xg = x.cuda().requires_grad_(True)
    logp = m(xg)
    assert logp.requires_grad and logp.shape == (B, T, 4)
    loss = torch.nn.functional.cross_entropy(logp.permute(0, 2, 1), y.cuda())
    loss.backward()
    ref_loss, ref_logp, ref_grads, ref_dx = lo.training_reference(params, h0, c0, x, y)
    _, _, grads64, dx64 = lo.training_reference(params, h0, c0, x, y, dtype=torch.float64)
    assert abs(float(loss.detach()) - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    assert (logp.detach().cpu() - ref_logp).abs().max() < 2e-5
    worst = worst64 = 0.0
    for name, p in m.named_parameters():
        e32, e64 = rel_err(p.grad.cpu(), ref_grads[name]), rel_err(p.grad.cpu().double(), grads64[name])
        worst = max(worst, min(e32, e64))
        worst64 = max(worst64, rel_err(ref_grads[name].double(), grads64[name]))
        assert min(e32, e64) < REL_TOL, (name, e32, e64)
    assert min(rel_err(xg.grad.cpu(), ref_dx), rel_err(xg.grad.cpu().double(), dx64)) < REL_TOL
    # eval mode on the same module = the inference kernels, same numbers as the training forward without dropout
    m.eval()
    with torch.no_grad():
        assert (m(x.cuda()) - logp.detach()).abs().max() < 2e-5
    print(f"B={B} T={T} F={F} H={H}: worst relative gradient error {worst:.2e} (nearer of the fp32 / float64 reference runs), those two runs differ by {worst64:.2e}")


def test_dropout_masks_and_an_optimizer_step(lib):
    """Dropout comes from torch's CUDA RNG (as in the reference); the masks it drew are recovered from the stream and replayed
    on the CPU reference.  Then one Adam step with gradient-norm clipping (main.py:130-135 / trainer gradient_clip_val)."""
    B, T, F, H = 4, 25, 44, 240
    m = make_model(1, F, B, H).cuda().train()
    params, h0, c0 = lo.reference_params(1, F, B, H)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, F, generator=g)
    y = torch.randint(0, 4, (B, T), generator=g)
    torch.manual_seed(123)
    logp = m(x.cuda())
    torch.manual_seed(123)                                     # replay: the forward draws exactly two [B,T,2H] masks, in order
    ones = torch.ones(B, T, 2 * H, device="cuda")
    masks = (m.dropout(ones).cpu(), m.dropout(ones).cpu())
    assert set(np.unique(masks[0].numpy()).round(4)) <= {0.0, 1.25}
    loss = torch.nn.functional.cross_entropy(logp.permute(0, 2, 1), y.cuda())
    loss.backward()
    ref_loss, _, ref_grads, _ = lo.training_reference(params, h0, c0, x, y, masks)
    assert abs(float(loss.detach()) - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    for name, p in m.named_parameters():
        assert rel_err(p.grad.cpu(), ref_grads[name]) < REL_TOL, name
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
    before = m.linear.weight.detach().clone()
    opt.step()
    assert not torch.equal(before, m.linear.weight)
    m.eval()                                                   # the inference path repacks the updated weights
    with torch.no_grad():
        out = m(x.cuda())
    new_params = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    assert (out.cpu() - lo.forward_torch(new_params, h0, c0, x)).abs().max() < 2e-5


def test_loss_decreases_over_a_few_steps(lib):
    B, T, F = 8, 200, 44
    m = make_model(3, F, B, 240).cuda().train()
    g = torch.Generator().manual_seed(3)
    y = torch.randint(0, 4, (B, 1), generator=g).expand(B, T).contiguous()
    x = torch.randn(B, T, F, generator=g) + 2.0 * torch.nn.functional.one_hot(y, F).float()    # the label is readable from x
    opt = torch.optim.Adam(m.parameters(), lr=0.01)
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(m(x.cuda()).permute(0, 2, 1), y.cuda())
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < 0.8 * losses[0], losses


def test_cluster_and_streaming_kernels_agree(lib, monkeypatch):
    """hidden_size 240 runs the cluster-resident kernels (lstm_train_cluster.cu); HSSB_TRAIN_IMPL=stream forces the generic
    ones (lstm_train.cu).  Same step on a batch with a ragged last row group, more steps than the CPU oracle affords."""
    B, T, F = 21, 300, 44
    g = torch.Generator().manual_seed(4)
    x = torch.randn(B, T, F, generator=g).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    results = []
    # cluster: fp32 cluster kernels (reduce-scatter backward); gather: all-gather backward; stream: generic kernels
    for impl in ("cluster", "gather", "stream"):
        monkeypatch.setenv("HSSB_TRAIN_IMPL", impl)
        m = make_model(11, F, B, 240).cuda().train()
        m.dropout.p = 0.0
        loss = torch.nn.functional.cross_entropy(m(x).permute(0, 2, 1), y)
        loss.backward()
        results.append((float(loss.detach()), {n: p.grad.clone() for n, p in m.named_parameters()}))
    for loss, grads in results[:2]:
        assert abs(loss - results[2][0]) < 1e-6 * abs(results[2][0])
        for name in grads:
            assert rel_err(grads[name], results[2][1][name]) < 2e-5, name


@pytest.mark.parametrize("B,T", [(21, 300), (50, 2000), (70, 130), (200, 260)])
def test_tensor_core_forward_agrees_with_the_fp32_kernels(lib, monkeypatch, B, T):
    """The default training forward of the reference geometry runs projection + recurrence on the tcgen05 kernels
    (hssb_lstm_train_forward_tc: split-fp16 operands, fp32 accumulation, approximate exp2 / reciprocal); HSSB_TRAIN_IMPL=stream
    is the plain fp32 path.  Shapes: the reference's training shape (50 x 2000, main.py:130) and batches that use 1, 2 and 3
    sub-tiles per cluster with ragged last groups.

    (a) Per layer, everything the forward hands to back-propagation -- activated gates, cell states, raw h, final states -- within
        2e-5 of the fp32 kernels' (values in [-1, 1], cells a few units).  The backward kernels are shared, so this is the whole
        difference between the two paths.
    (b) The full step (dropout off, so that (a)'s tensors are the step's): loss within 1e-5 relative, log-probabilities within
        5e-5; every gradient within 2e-4 of its tensor's max-abs when no ReLU input changes sign between the two forwards,
        else within 1 % of its norm: of the ~10^7 ReLU inputs a handful lie within the 1e-6 the forwards differ by, their units
        switch sides and each moves the gradient by one sample's worth (the count is printed)."""
    from hss.model import _train

    F = 44
    g = torch.Generator().manual_seed(B)
    x = (3.0 * torch.randn(B, T, F, generator=g)).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    m = make_model(13, F, B, 240).cuda().train()
    h0, c0 = m.h0.cuda().float(), m.c0.cuda().float()
    captured = {}
    real_save = torch.autograd.function.FunctionCtx.save_for_backward

    def layer_outputs(packed):
        outs = []
        with torch.no_grad():
            inp, h, c = x, h0, c0
            for li, lstm in enumerate((m.lstm_1, m.lstm_2)):
                ctx = type("Ctx", (), {"save_for_backward": lambda self, *t: captured.__setitem__("saved", t)})()
                args = [getattr(lstm, f"{k}_l0{sfx}") for sfx in ("", "_reverse") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
                out, hn, cn = _train.BiLSTMLayerFunction.forward(ctx, inp, h, c, *args, None if packed is None else packed[li])
                gates, cells = captured["saved"][7], captured["saved"][8]
                outs.append((gates.clone(), cells.clone(), out.clone(), hn.clone(), cn.clone()))
                inp, h, c = torch.relu(out), hn, cn
        return outs

    monkeypatch.setenv("HSSB_TRAIN_IMPL", "tc")
    packed = _train._packed_layers(m, x.device)
    assert packed[0] is not None
    tc, ref = layer_outputs(packed), layer_outputs(None)
    for li in range(2):
        for name, a, b in zip(("gates", "cells", "out", "hn", "cn"), tc[li], ref[li]):
            d = float((a - b).abs().max())
            assert d < 2e-5, (li, name, d)
    flips = sum(int(((tc[li][2] > 0) != (ref[li][2] > 0)).sum()) for li in range(2))

    results = []
    for impl in ("tc", "stream"):
        monkeypatch.setenv("HSSB_TRAIN_IMPL", impl)
        m = make_model(13, F, B, 240).cuda().train()
        m.dropout.p = 0.0
        loss, logp = m.training_loss(x, y)
        loss.backward()
        results.append((float(loss.detach()), logp.detach(), {n: p.grad.clone() for n, p in m.named_parameters()}))
    (l_tc, logp_tc, g_tc), (l_ref, logp_ref, g_ref) = results
    assert abs(l_tc - l_ref) < 1e-5 * abs(l_ref)
    assert (logp_tc - logp_ref).abs().max() < 5e-5
    worst = max(float((g_tc[n] - g_ref[n]).norm() / g_ref[n].norm()) for n in g_ref)
    worst_max = max(rel_err(g_tc[n], g_ref[n]) for n in g_ref)
    print(f"B={B} T={T}: loss {l_tc:.7f} vs {l_ref:.7f}, {flips} ReLU inputs of {2 * B * T * 480} change sign, "
          f"gradient difference: {worst:.2e} of the tensor norm, {worst_max:.2e} of its max-abs")
    assert worst_max < REL_TOL if flips == 0 else worst < 1e-2


def test_tf32_split_gemms_match_fp32(lib, monkeypatch):
    """Back-propagation's GEMMs (dG^T x, dG^T h_prev, dG W_ih) run as three TF32 tensor-core GEMMs on operands split by
    hssb_split_tf32; HSSB_TRAIN_GEMM=fp32 runs torch's SIMT fp32 ones.  The split itself: hi + lo == a exactly, hi has 13 zero
    low bits.  Gradients of the same step both ways within 2e-5 of each tensor's max-abs (sums of 10^4..10^5 products)."""
    from hss.model import _train

    g = torch.Generator().manual_seed(2)
    a = (torch.randn(100003, generator=g) * torch.logspace(-30, 30, 100003)).cuda()
    a[5], a[6], a[7] = float("inf"), 0.0, -1e-42            # inf passes through, zero and subnormals are fine
    hi, lo = _train._split_tf32(a)
    assert torch.equal(hi + lo, a)
    assert int((hi[torch.isfinite(hi)].view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float((lo[torch.isfinite(a)].abs() / a[torch.isfinite(a)].abs().clamp(min=1e-37)).max()) <= 2.0 ** -11

    B, T, F = 21, 300, 44
    x = torch.randn(B, T, F, generator=g).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    results = []
    for mode in ("tf32x3", "fp32"):
        monkeypatch.setenv("HSSB_TRAIN_GEMM", mode)
        m = make_model(17, F, B, 240).cuda().train()
        m.dropout.p = 0.0
        xg = x.clone().requires_grad_(True)
        loss, _ = m.training_loss(xg, y)
        loss.backward()
        results.append(({n: p.grad.clone() for n, p in m.named_parameters()}, xg.grad.clone()))
    assert not torch.backends.cuda.matmul.allow_tf32 or True     # (the scope restores the caller's setting)
    worst = max(rel_err(results[0][0][n], results[1][0][n]) for n in results[1][0])
    worst = max(worst, rel_err(results[0][1], results[1][1]))
    print(f"tf32x3 vs fp32 GEMMs: worst relative gradient difference {worst:.2e}")
    assert worst < 2e-5


@pytest.mark.parametrize("B,T,nb", [(5, 40, 0), (21, 300, 0), (21, 300, 16), (21, 300, 32), (50, 500, 0), (70, 130, 0), (130, 90, 0), (300, 33, 0)])
def test_tensor_core_backward_agrees_with_the_fp32_kernel(lib, monkeypatch, B, T, nb):
    """Back-propagation through time on the tcgen05 kernel (hssb_lstm_train_backward_tc; the default after a tensor-core forward)
    against the fp32 cluster kernel (HSSB_TRAIN_BWD=cluster) on the SAME forward.
    (a) Both with the plain fp32 gradient GEMMs, so that the recurrence kernels are the only difference: every parameter gradient
        and the input gradient within 2e-5 of the tensor's max-abs (the recurrent product runs on split-fp16 operands scaled from
        max|d_out|; measured 4e-7 .. 9e-7).
    (b) The default path -- K5b's own TF32-split outputs in the [row][direction][gate] layout, its bias-gradient sums, the merged
        three-pass TF32 GEMMs -- against (a)'s fp32 reference: within 1e-4 (different GEMM shapes, i.e. another summation order
        over up to 10^5 products).
    Batches cover 8 / 16 / 32 columns per cluster (forced or chosen), ragged last groups and more groups than one wave."""
    F = 44
    g = torch.Generator().manual_seed(B + T)
    x = torch.randn(B, T, F, generator=g).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    if nb:
        monkeypatch.setenv("HSSB_BPTT_NB", str(nb))
    results = []
    for impl, gemm in (("tc", "fp32"), ("cluster", "fp32"), ("tc", "tf32x3")):
        monkeypatch.setenv("HSSB_TRAIN_BWD", impl)
        monkeypatch.setenv("HSSB_TRAIN_GEMM", gemm)
        m = make_model(19, F, B, 240).cuda().train()
        m.dropout.p = 0.0
        xg = x.clone().requires_grad_(True)
        loss, _ = m.training_loss(xg, y)
        loss.backward()
        results.append(({n: p.grad.clone() for n, p in m.named_parameters()}, xg.grad.clone()))

    def worst_of(a, b):
        return max(max(rel_err(a[0][n], b[0][n]) for n in b[0]), rel_err(a[1], b[1]))

    worst, worst_default = worst_of(results[0], results[1]), worst_of(results[2], results[1])
    print(f"B={B} T={T} nb={nb or 'auto'}: worst relative gradient difference, tensor-core vs fp32 backward: {worst:.2e}; "
          f"default path (split outputs, merged TF32 GEMMs) vs fp32: {worst_default:.2e}")
    assert worst < 2e-5
    assert worst_default < 1e-4


def test_tensor_core_backward_keeps_tiny_and_huge_gradients(lib, monkeypatch):
    """The power-of-two scale follows max|d_out|: the same step with the loss multiplied by 1e-20 and by 1e+20 gives gradients
    scaled by exactly those factors (within the fp32 rounding of the factor itself)."""
    B, T, F = 9, 64, 44
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, T, F, generator=g).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    grads = []
    for factor in (1.0, 1e-20, 1e20):
        m = make_model(23, F, B, 240).cuda().train()
        m.dropout.p = 0.0
        loss, _ = m.training_loss(x, y)
        (loss * factor).backward()
        grads.append({n: p.grad.double() / factor for n, p in m.named_parameters()})
    for other in grads[1:]:
        for n in grads[0]:
            assert rel_err(other[n], grads[0][n]) < 1e-5, n


def test_repacked_weights_follow_the_optimizer(lib):
    """hssb_model_update: after every optimiser step the tcgen05 operands are re-packed in place; the training forward and the
    eval forward of the updated module agree with each other and the handle is reused, not recreated."""
    B, T, F = 6, 150, 44
    m = make_model(5, F, B, 240).cuda().train()
    m.dropout.p = 0.0
    from hss.optim import ClipAdam

    opt = ClipAdam(m.parameters(), lr=0.01, max_norm=1.0)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, T, F, generator=g).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    handles = set()
    for _ in range(3):
        opt.zero_grad()
        loss, logp = m.training_loss(x, y)
        handles.add(m._handle.value)
        loss.backward()
        opt.step()
        m.eval()
        with torch.no_grad():
            after = m(x)
        m.train()
        _, logp_after = m.training_loss(x, y)
        assert (after - logp_after).abs().max() < 2e-5
        assert (after - logp).abs().max() > 1e-4          # the step changed the outputs
    assert len(handles) == 1


def test_fused_head_loss_and_clip_adam_match_the_eager_ops(lib):
    """hssb_ce_head_forward / _backward (linear + log_softmax + CrossEntropyLoss of segmenter.py:86-87 + main.py:69-70) and
    hssb_clip_adam_step (clip_grad_norm_ 1.0 + Adam lr 0.01 * 0.9^epoch, main.py:130-135,226) against the eager torch ops the
    reference runs, over three steps of two epochs on two identically initialised modules (dropout off: p = 0)."""
    from hss.optim import ClipAdam

    B, T, F = 6, 40, 44
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, T, F, generator=g).cuda()
    y = torch.randint(0, 4, (B, T), generator=g).cuda()
    a = make_model(5, F, B, 240).cuda().train()
    b = make_model(5, F, B, 240).cuda().train()
    a.dropout.p = b.dropout.p = 0.0
    opt_a = torch.optim.Adam(a.parameters(), lr=0.01)
    sched = torch.optim.lr_scheduler.LambdaLR(opt_a, lr_lambda=lambda epoch: 0.9 ** epoch)
    opt_b = ClipAdam(b.parameters(), lr=0.01, max_norm=1.0)
    for step in range(3):
        opt_a.zero_grad()
        logp_a = a(x)
        loss_a = torch.nn.functional.cross_entropy(logp_a.permute(0, 2, 1), y)
        loss_a.backward()
        raw = [p.grad.detach().clone() for p in a.parameters()]
        norm_a = torch.nn.utils.clip_grad_norm_(a.parameters(), 1.0)
        opt_b.zero_grad()
        loss_b, logp_b = b.training_loss(x, y)
        loss_b.backward()
        assert abs(float(loss_a.detach()) - float(loss_b.detach())) < 2e-6 * max(1.0, abs(float(loss_a.detach())))
        assert (logp_a.detach() - logp_b).abs().max() < 2e-6
        for (name, _), pb, g in zip(a.named_parameters(), b.parameters(), raw):
            assert rel_err(pb.grad, g) < 2e-5, (step, name)                 # fused head + loss backward == the eager chain
            pb.grad.copy_(g)     # Adam divides by sqrt(v): entries whose gradient is rounding noise would amplify it -- feed both
            #                      optimisers the SAME gradients so that the comparison below tests the optimiser kernel alone
        opt_a.step()
        opt_b.step()
        assert abs(float(opt_b.grad_norm) - float(norm_a)) < 1e-5 * float(norm_a)
        for (name, pa), pb in zip(a.named_parameters(), b.parameters()):
            assert float((pb.detach() - pa.detach()).abs().max()) < 2e-6 * opt_b.current_lr + 1e-7 * float(pa.detach().abs().max()), (step, name)
        if step == 1:                       # epoch boundary: lr 0.01 -> 0.009
            sched.step()
            opt_b.set_epoch(1)
            assert abs(opt_b.current_lr - opt_a.param_groups[0]["lr"]) < 1e-12
    b.eval()
    with torch.no_grad():
        out = b(x)                          # the inference kernels see the updated (repacked) parameters
    a.eval()
    with torch.no_grad():
        assert (out - a(x)).abs().max() < 1e-4
