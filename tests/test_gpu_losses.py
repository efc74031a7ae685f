"""GPU parity, K2: fused focal / smooth-L1 forward+backward vs the numpy oracle (fp32 restatement of
model/losses.py).  Tolerance: north_star's 1e-5 relative for losses; gradients 1e-5 relative with a
1e-7 absolute floor (values are ~1e-3/n_pos)."""
import numpy as np
import pytest
import torch

from oracle import losses_np as OL

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def make_case(seed, B, N, C, p_ignore=0.1, p_pos=0.1, logit_mu=-2.0):
    rs = np.random.RandomState(seed)
    state = rs.choice([-1, 0, 1], (B, N), p=[p_ignore, 1 - p_ignore - p_pos, p_pos]).astype(np.float32)
    y_cls = np.zeros((B, N, C + 1), np.float32)
    y_cls[:, :, -1] = state
    hot = rs.randint(0, C, (B, N))
    bb, nn = np.nonzero(state == 1)
    y_cls[bb, nn, hot[bb, nn]] = 1
    p = (1 / (1 + np.exp(-rs.normal(logit_mu, 2, (B, N, C))))).astype(np.float32)
    y_reg = np.concatenate([rs.normal(0, 1, (B, N, 4)).astype(np.float32), state[:, :, None]], axis=2)
    r = (y_reg[:, :, :4] + rs.normal(0, 0.2, (B, N, 4))).astype(np.float32)
    return y_cls, p, y_reg, r


def close(got, want, rtol=RTOL, atol=0.0):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    return np.all(np.abs(got - want) <= atol + rtol * np.abs(want))


@pytest.mark.parametrize("C", [1, 3, 80])
@pytest.mark.parametrize("bce", ["tf2", "logits"])
def test_focal_vs_oracle(rn, C, bce):
    y_cls, p, _, _ = make_case(10 + C, 2, 3001, C)
    p[0, 0, 0], p[0, 1, 0], p[0, 2, C - 1], p[0, 3, 0] = 0.0, 1.0, 1e-8, 1 - 1e-8
    want_l, want_g = OL.focal(bce=bce)(y_cls, p, return_grad=True)
    yp = torch.tensor(p, device="cuda", requires_grad=True)
    loss = rn.focal(bce=bce)(torch.tensor(y_cls, device="cuda"), yp)
    loss.backward()
    assert close(loss.item(), want_l)
    assert close(yp.grad.cpu().numpy(), want_g, atol=1e-7 * float(np.abs(want_g).max()))
    # ignored anchors get exactly zero gradient
    assert not yp.grad.cpu().numpy()[y_cls[:, :, -1] == -1].any()


def test_smooth_l1_vs_oracle(rn):
    _, _, y_reg, r = make_case(20, 3, 2500, 1)
    r[0, 5] = y_reg[0, 5, :4]
    y_reg[0, 5, 4] = 1
    want_l, want_g = OL.smooth_l1()(y_reg, r, return_grad=True)
    yp = torch.tensor(r, device="cuda", requires_grad=True)
    loss = rn.smooth_l1()(torch.tensor(y_reg, device="cuda"), yp)
    (2.0 * loss).backward()                      # upstream gradient is applied in backward
    assert close(loss.item(), want_l)
    assert close(yp.grad.cpu().numpy(), 2.0 * want_g, atol=1e-9)
    assert not yp.grad.cpu().numpy()[y_reg[:, :, 4] != 1].any()
    for sigma in (1.0, 2.5):
        wl = OL.smooth_l1(sigma)(y_reg, r)
        assert close(rn.smooth_l1(sigma)(y_reg, r), wl)        # numpy in -> numpy scalar out


def test_golden_tf_half(rn, golden_tf):
    g = golden_tf
    for mode in ("tf2", "logits"):
        yp = torch.tensor(g['loss_p'], device="cuda", requires_grad=True)
        l = rn.focal(bce=mode)(torch.tensor(g['loss_y_cls'], device="cuda"), yp)
        l.backward()
        assert close(l.item(), g['focal_%s_loss' % mode])
        assert close(yp.grad.cpu().numpy(), g['focal_%s_grad' % mode], atol=1e-7 * float(np.abs(g['focal_%s_grad' % mode]).max()))
    yp = torch.tensor(g['loss_r'], device="cuda", requires_grad=True)
    l = rn.smooth_l1()(torch.tensor(g['loss_y_reg'], device="cuda"), yp)
    l.backward()
    assert close(l.item(), g['sl1_loss']) and close(yp.grad.cpu().numpy(), g['sl1_grad'], atol=1e-9)


@pytest.mark.parametrize("C", [1, 5])
def test_fused_equals_separate_and_normalizer(rn, C):
    y_cls, p, y_reg, r = make_case(30 + C, 2, 4099, C)
    t = lambda a: torch.tensor(a, device="cuda")
    losses, gc, gr = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p))
    wf, wgf = OL.focal()(y_cls, p, return_grad=True)
    ws, wgs = OL.smooth_l1()(y_reg, r, return_grad=True)
    npos = float((y_cls[:, :, -1] == 1).sum())
    l = losses.cpu().numpy()
    assert close(l[0], wf) and close(l[1], ws) and l[2] == max(1.0, npos)
    assert close(gc.cpu().numpy(), wgf, atol=1e-7 * float(np.abs(wgf).max())) and close(gr.cpu().numpy(), wgs, atol=1e-9)
    # explicit (global) normaliser: a device tensor of per-page counts, e.g. the K1 by-product all-reduced
    counts = torch.tensor([1000, 500], dtype=torch.int32, device="cuda")
    losses2, gc2, _ = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p), normalizer=counts)
    wf2, wgf2 = OL.focal()(y_cls, p, return_grad=True, normalizer=1500.0)
    assert close(losses2.cpu().numpy()[0], wf2) and close(gc2.cpu().numpy(), wgf2, atol=1e-7 * float(np.abs(wgf2).max()))
    # autograd wrapper
    cp, rp = t(p).requires_grad_(), t(r).requires_grad_()
    lf, ls = rn.detection_loss(t(y_reg), t(y_cls), rp, cp)
    (lf + ls).backward()
    assert close(cp.grad.cpu().numpy(), wgf, atol=1e-7 * float(np.abs(wgf).max())) and close(rp.grad.cpu().numpy(), wgs, atol=1e-9)


def test_shared_state_flag_is_identical_for_k1_targets(rn):
    """RN_LOSS_SHARED_STATE: smooth-L1 reads the state from the label rows.  For targets produced by K1 the two
    state columns are the same, so the gradients must be bit-identical to the default mode; the loss sums are
    accumulated in a different order by the C=1 fast kernel (1e-6 relative)."""
    import synthetic
    hw = (256, 320)
    anchors = rn.anchors_for_shape(hw + (3,))
    imgs = [synthetic.PageShape(hw + (3,)), synthetic.PageShape((256, 300, 3))]
    anns = [synthetic.gt_for_page(2, i, hw=hw, gmax=6) for i in range(2)]
    y_reg, y_cls, npos = rn.anchor_targets_bbox(anchors, imgs, anns, 1, output="torch", return_npos=True)
    assert torch.equal(y_reg[..., -1], y_cls[..., -1])
    cls, reg = synthetic.training_predictions(2, 2, anchors.shape[0], classes=1)
    t = lambda a: torch.tensor(a, device="cuda")
    a = rn.detection_losses(y_reg, y_cls, t(reg), t(cls), normalizer=npos)
    b = rn.detection_losses(y_reg, y_cls, t(reg), t(cls), normalizer=npos, shared_state=True)
    assert torch.allclose(a[0], b[0], rtol=1e-6, atol=0) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    step = rn.pipeline.TargetLossStep(hw + (3,), 2, 8, 1)
    step.load_annotations(imgs, anns)
    step.load_predictions(t(cls), t(reg))
    step.run()
    assert torch.equal(step.losses, b[0]) and torch.equal(step.grad_cls, a[1]) and torch.equal(step.grad_reg, a[2])
    assert torch.equal(step.y_reg, y_reg) and torch.equal(step.y_cls, y_cls)


@pytest.mark.parametrize("N", [1, 2, 63, 64, 65, 4099, 100001, 303104 * 2 + 7])
def test_c1_fast_kernel_vs_oracle(rn, N):
    """The C=1 fused kernel (k_loss_c1_fast: both losses, shared state, gamma 2, TF2 cross-entropy) at row
    counts around its warp-chunk / CTA-range boundaries, with probabilities on and outside the clip range,
    a soft label (general expression, out of line) and positives next to chunk edges."""
    y_cls, p, y_reg, r = make_case(60 + N % 97, 1, N, 1, p_ignore=0.05, p_pos=0.03, logit_mu=-3.0)
    y_reg[:, :, 4] = y_cls[:, :, 1]                                  # K1 invariant: both state columns agree
    for k, v in enumerate((0.0, 1.0, 1e-8, 1 - 1e-8)):
        if k < N:
            p[0, k, 0] = v
    for k in (0, 31, 32, 63, 64, N - 1):
        if 0 <= k < N:
            y_cls[0, k] = (1.0, 1.0)
            y_reg[0, k, 4] = 1.0
    if N > 70:
        y_cls[0, 70] = (0.3, 0.0)                                    # soft label on a non-ignored row
        y_reg[0, 70, 4] = 0.0
    t = lambda a: torch.tensor(a, device="cuda")
    losses, gc, gr = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p), shared_state=True)
    ref = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p))       # generic kernel, separate state columns
    wf, wgf = OL.focal()(y_cls, p, return_grad=True)
    ws, wgs = OL.smooth_l1()(y_reg, r, return_grad=True)
    l = losses.cpu().numpy()
    assert close(l[0], wf) and close(l[1], ws) and l[2] == max(1.0, float((y_cls[:, :, 1] == 1).sum()))
    assert close(gc.cpu().numpy(), wgf, atol=1e-7 * float(np.abs(wgf).max())) and close(gr.cpu().numpy(), wgs, atol=1e-9)
    assert torch.equal(gc, ref[1]) and torch.equal(gr, ref[2])       # element-wise arithmetic is the same


def test_no_positive_and_nan_in_ignored_rows(rn):
    """Normaliser is max(1, 0) = 1 with no positives; NaN predictions in ignored / non-positive rows do
    not leak (TF gathers those rows away before any arithmetic)."""
    y_cls, p, y_reg, r = make_case(40, 1, 1000, 1, p_pos=0.0)
    ign = y_cls[0, :, -1] == -1
    p[0, ign, 0] = np.nan
    r[0, :, :] = np.nan
    t = lambda a: torch.tensor(a, device="cuda")
    losses, gc, gr = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p))
    l = losses.cpu().numpy()
    pm = np.where(ign[None, :, None], np.float32(0.5), p)
    assert l[2] == 1.0 and l[1] == 0.0 and close(l[0], OL.focal()(y_cls, pm))
    assert np.isfinite(gc.cpu().numpy()).all() and not gr.cpu().numpy().any()


def test_full_size_properties(rn):
    """Config-2 size (16 x 200,700 anchors): the loss is a sum over pages -> compare the full-batch result
    with per-page launches sharing the same normaliser (linearity), and with the oracle on one page."""
    B, N = 16, 200700
    y_cls, p, y_reg, r = make_case(50, B, N, 1, p_ignore=0.01, p_pos=0.002, logit_mu=-4.6)
    t = lambda a: torch.tensor(a, device="cuda")
    Y_reg, Y_cls, R, P = t(y_reg), t(y_cls), t(r), t(p)
    losses, gc, gr = rn.detection_losses(Y_reg, Y_cls, R, P)
    total = losses.cpu().numpy()
    norm = torch.tensor([total[2]], device="cuda")
    acc = np.zeros(2)
    for b in range(B):
        lb, gcb, grb = rn.detection_losses(Y_reg[b:b + 1], Y_cls[b:b + 1], R[b:b + 1], P[b:b + 1], normalizer=norm)
        acc += lb.cpu().numpy()[:2].astype(np.float64)
        assert torch.equal(gcb[0], gc[b]) and torch.equal(grb[0], gr[b])
    assert close(total[:2], acc, rtol=1e-5)
    wl, wg = OL.focal()(y_cls[:1], p[:1], return_grad=True, normalizer=float(total[2]))
    assert close(gc[0].cpu().numpy(), wg[0], atol=1e-7 * float(np.abs(wg).max()))


@pytest.mark.parametrize("from_logits", [True, False])
@pytest.mark.parametrize("rows", [[150, 37, 10, 3, 1], [64, 64], [1], [33, 31, 65, 2]])
def test_per_level_heads_fused_sigmoid(rn, rows, from_logits):
    """N2: per-level (B, n_l, 1) logits / (B, n_l, 4) regression straight into the fused losses -- same losses and
    gradients as sigmoid + Concatenate(axis=1) + the reference losses (oracle), gradients w.r.t. the logits.
    Level sizes are chosen so that warp chunks (64 rows) cross level and page boundaries at every phase."""
    B, N = 3, sum(rows)
    y_cls, p, y_reg, r = make_case(70 + N, B, N, 1, p_ignore=0.05, p_pos=0.05, logit_mu=-3.0)
    y_reg[:, :, 4] = y_cls[:, :, 1]
    rs = np.random.RandomState(N)
    z = rs.normal(-3.0, 2.5, (B, N, 1)).astype(np.float32)
    z[0, 0, 0], z[0, min(1, N - 1), 0] = 40.0, -40.0                      # saturated sigmoid on both sides
    prob = (np.float32(1) / (np.float32(1) + np.exp(-z))).astype(np.float32) if from_logits else p
    want_f, want_gf = OL.focal()(y_cls, prob, return_grad=True)
    want_s, want_gs = OL.smooth_l1()(y_reg, r, return_grad=True)
    if from_logits:
        want_gf = want_gf * (prob * (np.float32(1) - prob))
    src = z if from_logits else p
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    bounds = np.concatenate([[0], np.cumsum(rows)])
    cls_l = [t(src[:, bounds[i]:bounds[i + 1]]) for i in range(len(rows))]
    reg_l = [t(r[:, bounds[i]:bounds[i + 1]]) for i in range(len(rows))]
    losses, g_cls, g_reg = rn.detection_losses_levels(t(y_reg), t(y_cls), reg_l, cls_l, from_logits=from_logits)
    l = losses.cpu().numpy()
    assert close(l[0], want_f) and close(l[1], want_s) and l[2] == max(1.0, float((y_cls[:, :, 1] == 1).sum()))
    got_gf = np.concatenate([g.cpu().numpy() for g in g_cls], axis=1)
    got_gs = np.concatenate([g.cpu().numpy() for g in g_reg], axis=1)
    assert close(got_gf, want_gf, atol=1e-7 * float(np.abs(want_gf).max()))
    assert close(got_gs, want_gs, atol=1e-9)
    if not from_logits:                                                   # probabilities in: bit-identical to the concatenated path
        ref = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p), shared_state=True)
        assert torch.equal(torch.cat(g_cls, 1), ref[1]) and torch.equal(torch.cat(g_reg, 1), ref[2])


@pytest.mark.parametrize("seed", list(range(10)))
def test_randomized_loss_parameters(rn, seed):
    """Random alpha / gamma / sigma / cross-entropy form / class count / batch shape, soft labels sprinkled in, through
    the fused entry point with and without the shared-state flag, the single-loss functors and an explicit
    normaliser -- every variant against the oracle (1e-5 relative; gradients + 1e-7 of their maximum)."""
    rs = np.random.RandomState(7000 + seed)
    B, N, C = int(rs.randint(1, 4)), int(rs.choice([1, 5, 130, 2049, 10007])), int(rs.choice([1, 1, 2, 4, 21]))
    alpha, gamma = float(rs.choice([0.25, 0.5, 0.9])), float(rs.choice([2.0, 2.0, 1.0, 1.5, 0.0]))
    sigma, bce = float(rs.choice([3.0, 1.0, 2.0])), str(rs.choice(["tf2", "logits"]))
    y_cls, p, y_reg, r = make_case(90 + seed, B, N, C, p_ignore=0.1, p_pos=0.15, logit_mu=float(rs.choice([-4.0, 0.0])))
    soft = rs.uniform(size=y_cls[:, :, :C].shape) < 0.02
    y_cls[:, :, :C] = np.where(soft, rs.uniform(0.05, 0.95, soft.shape).astype(np.float32), y_cls[:, :, :C])
    y_reg[:, :, 4] = y_cls[:, :, C]
    norm = None if seed % 2 else float(rs.randint(1, 500))
    wf, wgf = OL.focal(alpha, gamma, bce=bce)(y_cls, p, return_grad=True, normalizer=norm)
    ws, wgs = OL.smooth_l1(sigma)(y_reg, r, return_grad=True, normalizer=norm)
    t = lambda a: torch.tensor(a, device="cuda")
    nt = None if norm is None else torch.tensor([norm], dtype=torch.float32, device="cuda")
    tol_f = dict(atol=1e-7 * float(np.abs(wgf).max()) + 1e-12)
    for shared in (False, True):
        losses, gc, gr = rn.detection_losses(t(y_reg), t(y_cls), t(r), t(p), normalizer=nt, alpha=alpha, gamma=gamma,
                                             sigma=sigma, bce=bce, shared_state=shared)
        l = losses.cpu().numpy()
        assert close(l[0], wf, atol=1e-9) and close(l[1], ws, atol=1e-9), (seed, shared, l, wf, ws)
        assert close(gc.cpu().numpy(), wgf, **tol_f) and close(gr.cpu().numpy(), wgs, atol=1e-9)
    yp = t(p).requires_grad_()
    lf = rn.focal(alpha, gamma, bce=bce)(t(y_cls), yp, normalizer=nt)
    lf.backward()
    assert close(lf.item(), wf, atol=1e-9) and close(yp.grad.cpu().numpy(), wgf, **tol_f)
    rp = t(r).requires_grad_()
    ls = rn.smooth_l1(sigma)(t(y_reg), rp, normalizer=nt)
    ls.backward()
    assert close(ls.item(), ws, atol=1e-9) and close(rp.grad.cpu().numpy(), wgs, atol=1e-9)
