"""GPU parity tests proper: every C-ABI entry point (through rlvi_b200.ops / the drop-in modules, which
call ONLY librlvi_b200.so for N-sized work) against the oracle on the same seeded inputs and against the
golden vectors produced by the unmodified reference (tests/golden/, oracle/make_golden.py).

Tolerances (BASELINE.json north_star): 1e-9 relative for FP64 statistics and epsilon, 1e-5 for FP32
posteriors and losses, identical selection masks.  Raw posteriors in the collapse regime
are compared at SURVEY.md H1's max(1e-9, 4 * 2^-53 / mean pi).
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import deep_ref, rlvi_np

pytestmark = pytest.mark.gpu

F64_TOL = 1e-9
F32_TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    from rlvi_b200 import _lib
    _lib.load()                      # fail loudly if the extension is missing
    assert torch.cuda.is_available()
    return torch.device("cuda", 0)


def cu(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t if dtype is None else t.to(dtype)


def relmax(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def pi_tol(ref_pi):
    return max(F64_TOL, 4 * 2.0 ** -53 / max(float(np.mean(ref_pi)), 1e-300))


def test_concurrent_streams_use_separate_contexts(dev):
    """include/rlvi_b200.h, Conventions: calls on ONE context must be stream-ordered; the binding keeps one
    context per (device, stream), so work issued from two torch streams may overlap on the GPU and still
    gives the bits of the same calls made alone on the default stream."""
    from rlvi_b200 import _lib, ops, synth
    n, d = 300001, 64
    data = []
    for seed in (11, 12):
        X, y, theta = synth.logistic_data(n, d, seed=seed)
        params = np.concatenate([[0.1 * seed], theta])
        data.append(tuple(cu(a, dev) for a in (X, y, params)))

    def step(X, y, params):
        _, e, _ = ops.loss(ops.LOSS_LOGISTIC_CE, X, params, y=y, intercept=True, want_losses=False, want_e=True)
        pi, res = ops.fixed_point(None, e_work=e)
        return pi, res, ops.weighted_moments(X, pi, y=y)

    alone = [step(*t) for t in data]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=dev) for _ in data]
    outs = [None, None]
    for rep in range(3):
        for k, (st, t) in enumerate(zip(streams, data)):
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                outs[k] = step(*t)
    torch.cuda.synchronize()
    handles = {st.cuda_stream for st in streams}
    assert len(handles) == 2 and all((dev.index, h) in _lib._contexts for h in handles)
    for (pi0, res0, mom0), (pi1, res1, mom1) in zip(alone, outs):
        assert torch.equal(pi0, pi1) and torch.equal(mom0, mom1)
        assert ops.read_result(res0) == ops.read_result(res1)


# ==================================================================================================
# E-step: fixed point
# ==================================================================================================
@pytest.mark.parametrize("tag", ["n40", "n1000", "n4096", "n16384"])
def test_update_weights_golden(dev, tag):
    from rlvi_b200 import ops, rlvi
    g = load_golden("update_weights_" + tag)
    pi = rlvi.update_weights(g["losses"])                      # NumPy in -> NumPy out
    assert isinstance(pi, np.ndarray) and pi.dtype == np.float64
    assert relmax(pi, g["pi"]) < pi_tol(g["pi"])
    assert relmax(pi / pi.sum(), g["pi"] / g["pi"].sum()) < F64_TOL
    # iteration count, eps and err against the oracle's trace of the same loop
    _, eps, k, err = rlvi_np.fixed_point_trace(g["losses"])
    _, res = ops.fixed_point(cu(g["losses"], dev))
    r = ops.read_result(res)
    assert r["iters"] == k
    assert abs(r["eps"] - eps) <= F64_TOL * abs(eps)
    assert abs(r["err"] - err) <= 1e-6 * err + 1e-15
    assert abs(r["sum_pi"] - g["pi"].sum()) <= pi_tol(g["pi"]) * g["pi"].sum()


def test_update_weights_tol_maxiter_negative_losses(dev):
    from rlvi_b200 import ops, rlvi
    g = load_golden("update_weights_neg")
    pi = rlvi.update_weights(g["losses"], float(g["tol"]), int(g["maxiter"]))
    assert relmax(pi, g["pi"]) < pi_tol(g["pi"])
    _, res = ops.fixed_point(cu(g["losses"], dev), tol=float(g["tol"]), maxiter=int(g["maxiter"]))
    r = ops.read_result(res)
    assert r["iters"] == int(g["maxiter"]) and r["converged"] == 0


@pytest.mark.parametrize("regime", ["collapse", "clean", "bimodal", "balanced"])
@pytest.mark.parametrize("tol", [1e-1, 1e-3, 1e-6])
def test_fixed_point_stop_decision_across_regimes(dev, regime, tol):
    """fixed_point.cu decides most stop tests of the persistent kernel from the bound |sum pi' - sum pi| / sqrt(N) <= err
    (sum-only passes) and repeats a pass in full when the bound proves nothing: the iteration count, eps and the
    reported err must be the oracle's in every regime -- the collapse regime where the bound is tight (rlvi.py:8-20 with
    mostly large losses), clean data where pi stays near 1, a bimodal loss vector, and one whose posteriors move in
    opposite directions so that sum pi barely changes while err is large (the bound is useless there)."""
    from rlvi_b200 import ops
    rng = np.random.default_rng(11)
    n = 200001
    if regime == "collapse":
        losses = 0.5 * rng.chisquare(4, size=n) + 2.0
    elif regime == "clean":
        losses = 0.02 * rng.chisquare(1, size=n) - 3.0
    elif regime == "bimodal":
        losses = np.where(rng.random(n) < 0.3, 6.0 + rng.normal(size=n), -1.0 + 0.3 * rng.normal(size=n))
    else:
        losses = np.where(np.arange(n) % 2 == 0, -2.9444, 2.9444) + 1e-3 * rng.normal(size=n)   # pi ~ 0.95 / 0.05 around eps ~ 0.5
    ref, eps, k, err = rlvi_np.fixed_point_trace(losses, tol=tol, maxiter=100)
    pi, res = ops.fixed_point(cu(losses, dev), tol=tol, maxiter=100)
    r = ops.read_result(res)
    assert r["iters"] == k
    assert abs(r["eps"] - eps) <= F64_TOL * abs(eps) + 4 * 2.0 ** -53       # eps = 1 - mean(pi) is quantised to ulp(1)
    assert abs(r["err"] - err) <= 1e-6 * err + 1e-15
    assert relmax(pi.cpu().numpy(), ref) < max(pi_tol(ref), 4 * 2.0 ** -53 / max(eps, 1e-300) * 1e-0 if regime == "clean" else 0.0)


@pytest.mark.parametrize("n", [1, 2, 3, 7, 1023, 1025, 300001])
def test_fixed_point_ragged_and_unaligned(dev, n):
    from rlvi_b200 import ops
    rng = np.random.default_rng(n)
    losses = 0.5 * rng.chisquare(1, size=n + 1)
    ref, eps, k, err = rlvi_np.fixed_point_trace(losses[1:])
    base = cu(losses, dev)
    pi, res = ops.fixed_point(base[1:])                        # 8-byte aligned only -> scalar path
    r = ops.read_result(res)
    assert r["iters"] == k
    assert relmax(pi.cpu().numpy(), ref) < pi_tol(ref)
    pi2, res2 = ops.fixed_point(base[1:].clone())              # 16-byte aligned -> vector path
    assert ops.read_result(res2)["iters"] == k
    assert relmax(pi2.cpu().numpy(), ref) < pi_tol(ref)


@pytest.mark.parametrize("n", [2, 513, 4098, 70001, 606209, 2500003])
@pytest.mark.parametrize("slots", ["-1", "0", "8"])
def test_fixed_point_streaming_modes_on_ragged_sizes(dev, n, slots, monkeypatch):
    """The bulk-copy ring (RLVI_FP_CACHE_SLOTS=-1; the default above 2^22 samples per GPU), no residency (0) and a partial
    resident head (8), forced onto sizes where warps own no trip at all, only a partial trip, or whole trips plus a ragged
    rest -- the partial last trip of a warp's segment goes through the ring as a shorter copy, and every pass arms the next
    pass's first copies, which the kernel must drain when the loop ends early."""
    from rlvi_b200 import ops
    monkeypatch.setenv("RLVI_FP_CACHE_SLOTS", slots)
    rng = np.random.default_rng(n)
    losses = 0.5 * rng.chisquare(1, size=n)
    ref, eps, k, err = rlvi_np.fixed_point_trace(losses)
    pi, res = ops.fixed_point(cu(losses, dev))
    r = ops.read_result(res)
    assert r["iters"] == k
    assert relmax(pi.cpu().numpy(), ref) < pi_tol(ref)
    # a stop after very few passes (tol large, then maxiter = 2): armed copies of a pass that never runs
    for kw in ({"tol": 0.5}, {"maxiter": 2}):
        ref2, _, k2, _ = rlvi_np.fixed_point_trace(losses, **kw)
        pi2, res2 = ops.fixed_point(cu(losses, dev), **kw)
        assert ops.read_result(res2)["iters"] == k2
        assert relmax(pi2.cpu().numpy(), ref2) < pi_tol(ref2)


def test_fixed_point_scale_and_precomputed_e(dev):
    from rlvi_b200 import ops
    rng = np.random.default_rng(5)
    r2 = rng.chisquare(1, size=50000) * 3.0
    s = 0.37
    ref, _, k, _ = rlvi_np.fixed_point_trace(s * r2)
    scale = torch.tensor([s], dtype=torch.float64, device=dev)
    pi, res = ops.fixed_point(cu(r2, dev), scale=scale)
    assert ops.read_result(res)["iters"] == k
    assert relmax(pi.cpu().numpy(), ref) < pi_tol(ref)
    e = cu(np.exp(-(s * r2)), dev)
    pi2, res2 = ops.fixed_point(None, e_work=e)
    assert ops.read_result(res2)["iters"] == k
    assert relmax(pi2.cpu().numpy(), ref) < pi_tol(ref)
    # in-place: pi_out aliases losses
    l = cu(s * r2, dev)
    pi3, _ = ops.fixed_point(l, out=l)
    assert relmax(pi3.cpu().numpy(), ref) < pi_tol(ref)


def test_fixed_point_eps_equals_one_gives_zero_weights(dev):
    """Quirk Q1: huge losses underflow mean(pi) so eps == 1.0, rho = inf, pi = 0 (the reference then
    divides by zero downstream; the kernel reproduces the all-zero posteriors)."""
    from rlvi_b200 import ops
    losses = np.full(4096, 800.0)
    with np.errstate(all="ignore"):
        ref = rlvi_np.update_weights(losses)
    pi, _ = ops.fixed_point(cu(losses, dev))
    out = pi.cpu().numpy()
    assert np.all(ref == 0.0) and np.all(out == 0.0)


def test_online_fixed_point_golden(dev):
    from rlvi_b200 import online
    g = load_golden("online_n100")
    res = online.cross_entropy(g["log_proba"], g["targets"])
    assert np.array_equal(res, g["residuals"])
    pi = online.update_weights_rlvi(g["residuals"])
    assert relmax(pi, g["pi"]) < F64_TOL
    rng = np.random.default_rng(3)
    for n in (1, 5, 100, 4097):
        l = rng.exponential(1.0, size=n)
        assert relmax(online.update_weights_rlvi(l), rlvi_np.update_weights_online(l)) < F64_TOL


def test_online_carry_over_extension(dev):
    """Opt-in extension: start a batch from the previous batch's mean posterior instead of 0.5."""
    from rlvi_b200 import online
    rng = np.random.default_rng(8)
    l1, l2 = rng.exponential(0.8, size=100), rng.exponential(0.8, size=100)

    def ref(losses, w0):      # online-learning/main.py:45-58 with weights = w0 instead of 0.5
        e = np.exp(-losses)
        w = np.full_like(losses, w0)
        for _ in range(100):
            avg = np.mean(w)
            rho = avg / (1 - avg)
            new = rho * e / (1 + rho * e)
            if np.linalg.norm(new - w) < 1e-3:
                break
            w = new.copy()
        return new / (np.max(new) * len(new)), np.mean(new)

    w1, avg1 = online.update_weights_rlvi(l1, return_avg=True)
    r1, ravg1 = ref(l1, 0.5)
    assert relmax(w1, r1) < F64_TOL and abs(avg1 - ravg1) < F64_TOL
    w2 = online.update_weights_rlvi(l2, init_weight=avg1)
    assert relmax(w2, ref(l2, ravg1)[0]) < F64_TOL
    assert relmax(w2, rlvi_np.update_weights_online(l2)) > 1e-6      # really different from the 0.5 restart


@pytest.mark.parametrize("n", [1, 31, 33, 1031, 4097, 70001])
def test_no_out_of_bounds_writes(dev, n):
    """compute-sanitizer is closed on this pool: every output array is a view with sentinel guard bands on both
    sides, and the guards must survive every kernel (ragged sizes, 16-byte aligned and 8-byte aligned views)."""
    from rlvi_b200 import ops
    G, SENT = 64, -7.25e300
    rng = np.random.default_rng(n)
    d = 64

    def guarded(count, dtype=torch.float64, off=0, sent=SENT):
        buf = torch.full((count + 2 * G + off,), sent, dtype=dtype, device=dev)
        return buf, buf[G + off:G + off + count]

    def intact(buf, count, off=0, sent=SENT):
        head, tail = buf[:G + off], buf[G + off + count:]
        return bool((head == sent).all()) and bool((tail == sent).all())

    X = cu(rng.normal(size=(n, d)), dev)
    y = cu((rng.random(n) < 0.5).astype(np.float64), dev)
    w = cu(rng.random(n), dev)
    params = cu(rng.normal(size=d + 1) / 8, dev)
    for off in (0, 1):
        bl, l = guarded(n, off=off)
        be, e = guarded(n, off=off)
        bs, ws = guarded(2)
        ops.loss(ops.LOSS_LOGISTIC_CE, X, params, y=y, intercept=True, weights=w, losses_out=l, e_out=e, wsum_out=ws)
        assert intact(bl, n, off) and intact(be, n, off) and intact(bs, 2)
        assert torch.isfinite(l).all() and torch.isfinite(e).all()
        bp, pi = guarded(n, off=off)
        bw, ework = guarded(n, off=off)
        br, res = guarded(5)
        ops.fixed_point(l, e_work=ework, out=pi, result=res)
        assert intact(bp, n, off) and intact(bw, n, off) and intact(br, 5)
        ops.fixed_point(None, e_work=e, out=pi, result=res, variant=ops.FP_ONLINE)
        assert intact(bp, n, off) and intact(be, n, off)
        bsh, psh = guarded(n, off=off)
        ops.shift_sum_e(e, 1.3, 0.4, pi_out=psh)
        ops.shift_sum(l, 0.2, 0.4, pi_out=psh)
        assert intact(bsh, n, off)
    nm = 2 + 2 * d + d * d
    bm, mom = guarded(nm)
    ops.weighted_moments(X, w, y=y, out=mom)
    ops.weighted_moments(X, w, out=mom, want_gram=False)
    ops.weighted_moments(X, w, out=mom, center=params[1:].contiguous())
    assert intact(bm, nm)
    bg, g = guarded(d + 1)
    ops.logistic_grad(X, y, w, params, out=g)
    assert intact(bg, d + 1)
    U = torch.triu(torch.eye(d, dtype=torch.float64, device=dev) * 2.0)
    gp = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), params[1:], U.reshape(-1)]).contiguous()
    bl2, l2 = guarded(n)
    ops.loss(ops.LOSS_GAUSSIAN, X, gp, losses_out=l2)
    assert intact(bl2, n) and torch.isfinite(l2).all()
    # FP32 deep path
    c = 100
    logits = torch.randn(n, c, device=dev)
    labels = torch.randint(0, c, (n,), device=dev)
    F32S = -3.5e30
    bres, resid = guarded(n, torch.float32, sent=F32S)
    bwt, wt = guarded(n, torch.float32, sent=F32S)
    wt.fill_(1.0)
    resid.zero_()
    r = ops.wce_fwd_bwd(logits, labels, wt, resid, want_per_sample=True, want_correct=True)
    assert intact(bres, n, sent=F32S) and intact(bwt, n, sent=F32S)
    bew, ew = guarded(n, torch.float32, sent=F32S)
    ops.fixed_point_deep(resid, wt, e_work=ew)
    ops.fn_threshold(wt, truncate=True)
    assert intact(bres, n, sent=F32S) and intact(bwt, n, sent=F32S) and intact(bew, n, sent=F32S)
    assert torch.isfinite(wt).all() and float(wt.max()) == 1.0


def test_c_abi_from_c(dev):
    """The C ABI driven from a plain C++ program (CUDA runtime only, no Python / torch on that side), checked
    against a scalar restatement of the reference arithmetic written in C (tests/c_abi/abi_smoke.cu)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c_abi", "abi_smoke")
    if not os.path.exists(exe):                 # normally built by __graft_entry__.build() and shipped in-tree
        import __graft_entry__ as g
        g.build()
    for n in ("50000", "1031"):
        r = subprocess.run([exe, n], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "C-ABI PARITY OK" in r.stdout


def test_c_abi_error_codes(dev):
    """Bad arguments come back as negative codes + message (never a crash, never a silent fallback)."""
    from rlvi_b200 import _lib, ops
    x = torch.zeros(4, dtype=torch.float64, device=dev)
    with pytest.raises(_lib.RlviError, match="maxiter"):
        ops.fixed_point(x, maxiter=0)
    with pytest.raises(_lib.RlviError, match="alias"):
        ops.fixed_point(x, e_work=x, out=x)
    with pytest.raises(_lib.RlviError, match="pi0"):
        ops.fixed_point(x, pi0=1.5)
    X = torch.zeros((8, 2000), dtype=torch.float64, device=dev)
    with pytest.raises(_lib.RlviError, match="1 <= d <= 1024"):
        ops.loss(ops.LOSS_PCA, X, torch.zeros(2000, dtype=torch.float64, device=dev))
    with pytest.raises(_lib.RlviError, match="needs y"):
        ops.loss(ops.LOSS_SQRES, X[:, :8].contiguous(), torch.zeros(8, dtype=torch.float64, device=dev))
    with pytest.raises(_lib.RlviError, match="power"):
        ops.weighted_moments(X[:, :8].contiguous(), torch.ones(8, dtype=torch.float64, device=dev), power=3)
    with pytest.raises(_lib.RlviError, match="classes"):
        ops.wce_fwd_bwd(torch.zeros((2, 2048), device=dev), torch.zeros(2, dtype=torch.int64, device=dev),
                        torch.ones(2, device=dev), torch.zeros(2, device=dev))
    with pytest.raises(TypeError):
        ops.fixed_point(torch.zeros(4, dtype=torch.float64))               # CPU tensor: no CPU path
    # lengths are checked on the host side (the C ABI sees raw pointers only)
    X8 = X[:, :8].contiguous()
    v8, v7 = torch.ones(8, dtype=torch.float64, device=dev), torch.ones(7, dtype=torch.float64, device=dev)
    with pytest.raises(ValueError, match="weights must have 8"):
        ops.weighted_moments(X8, v7)
    with pytest.raises(ValueError, match="params must have 9"):
        ops.loss(ops.LOSS_LOGISTIC_CE, X8, v8, y=v8, intercept=True)
    with pytest.raises(ValueError, match="params must have 73"):
        ops.loss(ops.LOSS_GAUSSIAN, X8, v8)
    with pytest.raises(ValueError, match="y must have 8"):
        ops.logistic_grad(X8, v7, v8, torch.zeros(9, dtype=torch.float64, device=dev))
    with pytest.raises(ValueError, match="out must have 4"):
        ops.fixed_point(x, out=v7)
    with pytest.raises(ValueError, match="labels must have 2"):
        ops.wce_fwd_bwd(torch.zeros((2, 10), device=dev), torch.zeros(3, dtype=torch.int64, device=dev),
                        torch.ones(2, device=dev), torch.zeros(2, device=dev))
    # the library is still healthy afterwards
    pi, _ = ops.fixed_point(torch.rand(100, dtype=torch.float64, device=dev))
    assert torch.isfinite(pi).all()


@pytest.mark.parametrize("tag", ["n50", "n3000"])
def test_update_weights_constrained_golden(dev, tag):
    from rlvi_b200 import rlvi
    g = load_golden("constrained_" + tag)
    pi = rlvi.update_weights_constrained(g["losses"], float(g["n_eff"]))
    # Brent's own accuracy bounds the agreement: the constraint is met to ~1e-8
    assert abs(pi.sum() - float(g["n_eff"])) < 1e-3 * float(g["n_eff"])
    assert relmax(pi, g["pi"]) < 1e-6


def test_shift_sum(dev):
    from rlvi_b200 import ops
    rng = np.random.default_rng(11)
    l = rng.chisquare(2, size=77777)
    s, c = 0.8, 0.43
    t = np.exp(-l + s)
    ref = t / (c + t)
    pi = torch.empty(l.size, dtype=torch.float64, device=dev)
    out = ops.shift_sum(cu(l, dev), s, c, pi_out=pi)
    assert abs(out.item() - ref.sum()) < 1e-12 * ref.sum()
    assert relmax(pi.cpu().numpy(), ref) < 1e-14
    # the exp-free form on e = exp(-l): objective only (fast reciprocal) and with the weights (IEEE division)
    e = cu(np.exp(-l), dev)
    for off in (0, 1):                                           # aligned (vector) and unaligned (scalar) paths
        ev = torch.cat([torch.zeros(off, dtype=torch.float64, device=dev), e])[off:]
        out2 = ops.shift_sum_e(ev, np.exp(s), c)
        assert abs(out2.item() - ref.sum()) < 1e-12 * ref.sum()
        pi2 = torch.empty(l.size + off, dtype=torch.float64, device=dev)[off:]
        out3 = ops.shift_sum_e(ev, np.exp(s), c, pi_out=pi2)
        assert abs(out3.item() - ref.sum()) < 1e-12 * ref.sum()
        assert relmax(pi2.cpu().numpy(), ref) < 1e-14


# ==================================================================================================
# per-sample losses
# ==================================================================================================
SHAPES = [(1, 1), (5, 2), (33, 3), (40, 10), (257, 31), (1000, 32), (1031, 64), (777, 65), (500, 100),
          (300, 512), (130, 1000), (64, 1024)]


@pytest.mark.parametrize("n,d", SHAPES)
def test_losses_against_oracle(dev, n, d):
    from rlvi_b200 import ops
    rng = np.random.default_rng(1000 * n + d)
    X = rng.normal(size=(n, d))
    y = (rng.random(n) < 0.5).astype(np.float64)
    theta = rng.normal(size=d) / np.sqrt(d)
    b = 0.3
    w = rng.random(n)
    Xd, yd, wd = cu(X, dev), cu(y, dev), cu(w, dev)
    Xa = np.hstack([np.ones((n, 1)), X])
    tb = np.concatenate([[b], theta])

    # logistic CE with and without intercept (utils.py:19-21)
    ref = rlvi_np.cross_entropy(Xa, tb, y)
    l, e, ws = ops.loss(ops.LOSS_LOGISTIC_CE, Xd, cu(tb, dev), y=yd, intercept=True, weights=wd, want_e=True)
    assert relmax(l.cpu().numpy(), ref) < 1e-12
    assert relmax(e.cpu().numpy(), np.exp(-ref)) < 1e-12
    ws = ws.cpu().numpy()
    assert abs(ws[0] - w @ ref) < 1e-12 * abs(w @ ref) and abs(ws[1] - w.sum()) < 1e-13 * w.sum()
    l2, _, _ = ops.loss(ops.LOSS_LOGISTIC_CE, Xd, cu(theta, dev), y=yd, intercept=False)
    assert relmax(l2.cpu().numpy(), rlvi_np.cross_entropy(X, theta, y)) < 1e-12

    # softplus (utils.py:62-64)
    l3, _, _ = ops.loss(ops.LOSS_SOFTPLUS, Xd, cu(tb, dev), intercept=True)
    assert relmax(l3.cpu().numpy(), rlvi_np.softplus_loss(X, tb)) < 1e-12

    # squared residual (rlvi.py:72), squared distance (rlvi.py:49), PCA reconstruction (utils.py:77-79)
    yr = X @ theta + rng.normal(size=n)
    l4, _, _ = ops.loss(ops.LOSS_SQRES, Xd, cu(theta, dev), y=cu(yr, dev))
    assert relmax(l4.cpu().numpy(), (yr - X @ theta) ** 2) < 1e-11
    mu = rng.normal(size=d)
    l5, _, _ = ops.loss(ops.LOSS_SQDIST, Xd, cu(mu, dev))
    assert relmax(l5.cpu().numpy(), np.linalg.norm(mu - X, axis=1) ** 2) < 1e-12
    v = theta / np.linalg.norm(theta)
    l6, _, _ = ops.loss(ops.LOSS_PCA, Xd, cu(v, dev))
    refp = rlvi_np.pca_losses(X, v)
    assert np.max(np.abs(l6.cpu().numpy() - refp)) < 1e-12 * np.max(np.sum(X ** 2, axis=1))


@pytest.mark.parametrize("n,d", [(300001, 64), (150000, 16), (70001, 128), (20000, 256)])
def test_losses_many_tiles_per_cta(dev, n, d):
    """TMA path with several ring rounds per consumer warp (n >> 148 SMs x 16 warps x 32 rows) and a ragged tail."""
    from rlvi_b200 import ops
    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d))
    y = (rng.random(n) < 0.5).astype(np.float64)
    w = rng.random(n)
    tb = np.concatenate([[0.2], rng.normal(size=d) / np.sqrt(d)])
    ref = rlvi_np.cross_entropy(np.hstack([np.ones((n, 1)), X]), tb, y)
    for rep in range(2):
        l, e, ws = ops.loss(ops.LOSS_LOGISTIC_CE, cu(X, dev), cu(tb, dev), y=cu(y, dev), intercept=True,
                            weights=cu(w, dev), want_e=True)
        assert relmax(l.cpu().numpy(), ref) < 1e-12
        assert relmax(e.cpu().numpy(), np.exp(-ref)) < 1e-12
        assert abs(ws[0].item() - w @ ref) < 1e-12 * abs(w @ ref)
    v = tb[1:] / np.linalg.norm(tb[1:])
    l6, _, _ = ops.loss(ops.LOSS_PCA, cu(X, dev), cu(v, dev))
    assert np.max(np.abs(l6.cpu().numpy() - rlvi_np.pca_losses(X, v))) < 1e-12 * np.max(np.sum(X ** 2, axis=1))


def test_losses_unaligned_rows(dev):
    from rlvi_b200 import ops
    rng = np.random.default_rng(2)
    n, d = 999, 64
    flat = rng.normal(size=n * d + 1)
    X = flat[1:].reshape(n, d)
    theta = rng.normal(size=d)
    base = cu(flat, dev)
    Xd = base[1:].view(n, d)                                   # 8-byte aligned only
    l, _, _ = ops.loss(ops.LOSS_PCA, Xd, cu(theta / np.linalg.norm(theta), dev))
    ref = rlvi_np.pca_losses(X, theta / np.linalg.norm(theta))
    assert np.max(np.abs(l.cpu().numpy() - ref)) < 1e-11 * np.max(np.sum(X ** 2, axis=1))


@pytest.mark.parametrize("n,d", [(50, 2), (999, 3), (2048, 16), (1500, 64), (700, 65), (300, 128), (300001, 64),
                                 (100000, 32), (70001, 48), (33, 16)])
def test_gaussian_loss_against_oracle(dev, n, d):
    from rlvi_b200 import ops, utils
    rng = np.random.default_rng(d)
    A = rng.normal(size=(d, d)) / np.sqrt(d)
    cov = A @ A.T + 0.5 * np.eye(d)
    mu = rng.normal(size=d)
    X = mu + rng.normal(size=(n, d)) @ np.linalg.cholesky(cov).T
    ref = rlvi_np.gaussian_losses(X, mu, cov)
    params = utils._gaussian_params(cu(mu, dev), cu(cov, dev))
    l, _, _ = ops.loss(ops.LOSS_GAUSSIAN, cu(X, dev), params)
    assert relmax(l.cpu().numpy(), ref) < 1e-10


# ==================================================================================================
# weighted M-step statistics
# ==================================================================================================
MOM_SHAPES = [(1, 1), (7, 2), (40, 10), (257, 31), (1000, 33), (63, 64), (64, 64), (65, 64), (4096 + 17, 64),
              (100000, 64), (777, 65), (500, 100), (300, 512), (150, 1000), (64, 16), (20001, 128), (5000, 256),
              (4096 + 17, 512), (2000, 1024), (3000, 130)]


@pytest.mark.parametrize("n,d", MOM_SHAPES)
@pytest.mark.parametrize("power", [1, 2])
def test_weighted_moments_against_oracle(dev, n, d, power):
    from rlvi_b200 import ops
    rng = np.random.default_rng(7 * n + d + power)
    X = rng.normal(size=(n, d)) + 0.3
    y = rng.normal(size=n)
    w = rng.random(n)
    we = w ** power
    ref = rlvi_np.weighted_moments(X, we, y)
    for use_y in (True, False):
        out = ops.weighted_moments(cu(X, dev), cu(w, dev), y=cu(y, dev) if use_y else None, power=power)
        m = {k: v.cpu().numpy() for k, v in ops.split_moments(out, d).items()}
        assert abs(m["S0"] - ref["S0"]) < 1e-12 * ref["S0"]
        assert relmax(m["S1"], X.T @ w) < 1e-12                 # first power always (header contract)
        assert relmax(m["G"], ref["G"]) < 1e-12
        assert np.array_equal(m["G"], m["G"].T)                 # exactly symmetric
        if use_y:
            assert relmax(m["Sy"], ref["Sy"]) < 1e-12
            assert abs(m["Swy"] - ref["Swy"]) < 1e-12 * max(abs(ref["Swy"]), np.abs(we * y).sum() * 1e-3)
        else:
            assert not m["Sy"].any() and m["Swy"] == 0.0


def test_weighted_moments_no_gram_and_determinism(dev):
    from rlvi_b200 import ops
    rng = np.random.default_rng(9)
    n, d = 50000, 64
    X, w = cu(rng.normal(size=(n, d)), dev), cu(rng.random(n), dev)
    a = ops.weighted_moments(X, w)
    b = ops.weighted_moments(X, w)
    assert torch.equal(a, b)                                    # fixed-order reductions: same bits
    c = ops.weighted_moments(X, w, want_gram=False)
    # different kernels, different (fixed) summation orders: agreement relative to the largest entry
    assert float((c[:2 + 2 * d] - a[:2 + 2 * d]).abs().max() / a[:2 + 2 * d].abs().max()) < 1e-13
    assert not c[2 + 2 * d:].any()


@pytest.mark.parametrize("n,d", [(300001, 64), (50001, 32), (9999, 256)])
def test_moments_without_gram_tma_path(dev, n, d):
    from rlvi_b200 import ops
    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d)) + 0.2
    y = rng.normal(size=n)
    w = rng.random(n)
    for power in (1, 2):
        for use_y in (True, False):
            out = ops.weighted_moments(cu(X, dev), cu(w, dev), y=cu(y, dev) if use_y else None, power=power,
                                       want_gram=False)
            m = {k: v.cpu().numpy() for k, v in ops.split_moments(out, d).items()}
            we = w ** power
            assert abs(m["S0"] - we.sum()) < 1e-12 * we.sum()
            assert relmax(m["S1"], X.T @ w) < 1e-12
            if use_y:
                assert relmax(m["Sy"], X.T @ (we * y)) < 1e-12
                assert abs(m["Swy"] - we @ y) < 1e-12 * np.abs(we * y).sum()
            else:
                assert not m["Sy"].any() and m["Swy"] == 0.0


def test_weighted_moments_linearity_large(dev):
    """Size-independent property at a size the oracle cannot reach quickly: moments are linear in w."""
    from rlvi_b200 import ops
    n, d = 1 << 21, 64
    g = torch.Generator(device=dev).manual_seed(0)
    X = torch.randn((n, d), generator=g, device=dev, dtype=torch.float64)
    w1 = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    w2 = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
    m1, m2, m12 = (ops.weighted_moments(X, w) for w in (w1, w2, w1 + w2))
    scale = m12.abs().max()
    assert float((m1 + m2 - m12).abs().max() / scale) < 1e-12
    # and against cuBLAS on the same device (torch fp64 reference of the same contraction)
    G = (X * (w1 + w2)[:, None]).T @ X
    assert float((ops.split_moments(m12, d)["G"] - G).abs().max() / G.abs().max()) < 1e-12


@pytest.mark.parametrize("n,d", [(40, 10), (1031, 64), (777, 65), (300, 512), (300001, 64), (70001, 128), (9999, 256),
                                 (33, 16)])
def test_logistic_grad_against_oracle(dev, n, d):
    from rlvi_b200 import ops
    rng = np.random.default_rng(n + d)
    X = rng.normal(size=(n, d))
    y = (rng.random(n) < 0.5).astype(np.float64)
    w = rng.random(n)
    tb = rng.normal(size=d + 1) / np.sqrt(d)
    Xa = np.hstack([np.ones((n, 1)), X])
    ref = Xa.T @ (w * (rlvi_np.sigmoid(Xa @ tb) - y))
    out = ops.logistic_grad(cu(X, dev), cu(y, dev), cu(w, dev), cu(tb, dev))
    assert relmax(out.cpu().numpy(), ref) < 1e-12


# ==================================================================================================
# drop-in outer loops against the reference's golden outputs
# ==================================================================================================
@pytest.mark.parametrize("tag", ["n100_d2", "n768_d64"])
def test_mean_golden(dev, tag):
    from rlvi_b200 import rlvi
    g = load_golden("mean_" + tag)
    assert relmax(rlvi.mean(g["X"]), g["theta"]) < F64_TOL


@pytest.mark.parametrize("tag", ["n40_d10", "n768_d64"])
def test_linear_regression_golden(dev, tag):
    from rlvi_b200 import rlvi
    g = load_golden("linreg_" + tag)
    assert relmax(rlvi.linear_regression(g["X"], g["y"]), g["theta"]) < F64_TOL


def test_linear_regression_rank_deficient_and_badly_scaled(dev):
    """rlvi.py:71 solves with lstsq (gelsd): a duplicated column gets the minimum-norm solution; columns whose
    scales differ by 1e7 are still full rank.  The Gram route must agree in both cases."""
    from rlvi_b200 import rlvi
    rng = np.random.default_rng(3)
    X = rng.normal(size=(200, 4))
    X = np.hstack([X, X[:, :1]])
    y = X @ np.array([1.0, -2.0, 0.5, 3.0, 1.0]) + 0.01 * rng.normal(size=200)
    got = rlvi.linear_regression(X, y)
    assert relmax(got, rlvi_np.linear_regression(X, y)) < 1e-6
    assert abs(got[0] - got[4]) < 1e-8 * abs(got[0])
    rng = np.random.default_rng(5)
    X = rng.normal(size=(300, 4)) * np.array([1.0, 1e-7, 1e3, 1.0])
    y = X @ np.array([1.0, 2e7, -3e-3, 0.5]) + 0.1 * rng.normal(size=300)
    ref = rlvi_np.linear_regression(X, y)
    assert np.max(np.abs(rlvi.linear_regression(X, y) - ref) / np.abs(ref)) < 1e-6


def test_mm_log_reg_and_logistic_regression_golden(dev):
    from rlvi_b200 import rlvi, utils
    g = load_golden("mm_log_reg_n768_d64")
    assert relmax(utils.sigmoid(g["x_sig"]), g["sig"]) < 1e-15
    theta, losses = utils.mm_log_reg(g["X"], g["y"], g["w"])
    assert relmax(theta, g["theta"]) < F64_TOL
    assert relmax(losses, g["losses"]) < F64_TOL
    Xa = np.hstack([np.ones((g["X"].shape[0], 1)), g["X"]])
    assert relmax(utils.cross_entropy(Xa, g["theta"], g["y"]), g["ce_at_theta"]) < 1e-12
    g2 = load_golden("logreg_mm_n1500_d8")
    assert relmax(rlvi.logistic_regression(g2["X"], g2["y"], mstep="mm"), g2["theta"]) < 1e-8


def test_clf_predict(dev):
    from rlvi_b200 import utils
    rng = np.random.default_rng(12)
    X = rng.normal(size=(500, 7))
    theta = rng.normal(size=8)
    Xa = np.hstack([np.ones((500, 1)), X])
    ref = np.array(rlvi_np.sigmoid(Xa @ theta) > 0.5, dtype=int)          # utils.py:24-29
    out = utils.clf_predict(X, theta)
    assert out.dtype.kind == "i" and np.array_equal(out, ref)
    ref2 = np.array(rlvi_np.sigmoid(X @ theta[1:]) > 0.5, dtype=int)
    assert np.array_equal(utils.clf_predict(X, theta[1:], augment=False), ref2)


def test_sklearn_log_reg_golden(dev):
    """liblinear is third-party and stops at tol = 1e-4: agreement is to liblinear's accuracy; the
    in-place normalisation of the caller's weights (Q3) and the label-independent loss are exact."""
    from rlvi_b200 import utils
    g = load_golden("sklearn_loss_n600_d5")
    w = g["w"].copy()
    theta, losses = utils.sklearn_log_reg(g["X"], g["y"], w)
    assert np.array_equal(w, g["w_after"])
    assert relmax(theta, g["theta"]) < 5e-3
    assert relmax(losses, rlvi_np.softplus_loss(g["X"], theta)) < 1e-12
    assert relmax(losses, g["losses"]) < 5e-3


@pytest.mark.parametrize("tag", ["n400_d2", "n768_d64"])
def test_pca_golden(dev, tag):
    from rlvi_b200 import rlvi, utils
    g = load_golden("pca_" + tag)
    theta, losses = utils.pca(g["X"], g["w"])
    assert relmax(theta, g["theta_mstep"]) < 1e-8
    assert relmax(losses, g["losses_mstep"]) < 1e-8
    _, l0 = utils.pca(g["X"], g["w"], g["theta_init"])
    assert relmax(l0, g["losses_init"]) < 1e-12
    assert relmax(rlvi.pca(g["X"], theta_init=g["theta_init"]), g["theta"]) < 1e-7


@pytest.mark.parametrize("tag", ["n50_d2", "n2048_d16"])
def test_covariance_golden(dev, tag):
    from rlvi_b200 import rlvi, utils
    g = load_golden("cov_" + tag)
    cov, losses = utils.covariance(g["X"], g["w"])
    assert relmax(cov, g["cov_mstep"]) < F64_TOL
    assert relmax(losses, g["losses_mstep"]) < F64_TOL
    out = rlvi.covariance(g["X"], float(g["eps"]))
    assert relmax(out, g["cov"]) < 1e-5       # bounded by Brent's accuracy in the constrained E-step (H4)


@pytest.mark.parametrize("n,d", [(50, 2), (999, 3), (5000, 64), (3000, 128), (2000, 100)])
def test_covariance_large_mean_is_centred(dev, n, d):
    """The reference centres before the contraction (utils.py:103-105): a mean of 1e6 with unit variance must not
    cost digits (G/S0 - mu mu^T would lose ~12)."""
    from rlvi_b200 import ops, utils
    rng = np.random.default_rng(n + d)
    X = 1e6 + rng.normal(size=(n, d)) @ (np.eye(d) + 0.3 * rng.normal(size=(d, d)) / np.sqrt(d))
    w = rng.random(n)
    ref_cov, ref_losses = rlvi_np.covariance_mstep(X, w)
    cov, losses = utils.covariance(X, w)
    assert relmax(cov, ref_cov) < 1e-8          # limited by the reference's own rounding of x - mu at 1e6
    assert relmax(losses, ref_losses) < 1e-6
    # the centred statistics themselves, against NumPy on the same centre
    c = X.mean(axis=0)
    out = ops.weighted_moments(cu(X, dev), cu(w, dev), y=cu(w, dev), center=cu(c, dev))
    m = {k: v.cpu().numpy() for k, v in ops.split_moments(out, d).items()}
    Xc = X - c
    assert relmax(m["G"], (Xc * w[:, None]).T @ Xc) < 1e-12
    assert relmax(m["S1"], Xc.T @ w) < 1e-9
    assert relmax(m["Sy"], Xc.T @ (w * w)) < 1e-9


def test_covariance_singular_raises(dev):
    from rlvi_b200 import utils
    X = np.ones((20, 3))
    with pytest.raises(ValueError, match="Singular covariance matrix"):
        utils.covariance(X, np.ones(20))


def test_cuda_tensor_in_tensor_out(dev):
    from rlvi_b200 import rlvi
    g = load_golden("linreg_n768_d64")
    theta = rlvi.linear_regression(cu(g["X"], dev), cu(g["y"], dev))
    assert isinstance(theta, torch.Tensor) and theta.is_cuda
    assert relmax(theta.cpu().numpy(), g["theta"]) < F64_TOL


# ==================================================================================================
# one E+M step through the host-buffer entry point (bench.py's e2e call)
# ==================================================================================================
@pytest.mark.parametrize("n", [1000, 200001])
def test_em_step_logistic_host(dev, n):
    from rlvi_b200 import ops, synth
    X, y, theta = synth.logistic_data(n, 64, seed=1)
    params = np.concatenate([[0.1], theta])
    ref = rlvi_np.em_step_logistic(X, y, params)
    out = ops.em_step_logistic_host(X, y, params)
    assert out["result"]["iters"] == ref["iters"]
    assert relmax(out["pi"], ref["pi"]) < pi_tol(ref["pi"])
    m = out["moments"]
    d = 64
    tol = pi_tol(ref["pi"])
    assert abs(m[0] - ref["S0"]) < tol * ref["S0"]
    assert relmax(m[2:2 + d], ref["S1"]) < max(tol, 1e-9) * 10
    assert relmax(m[2 + 2 * d:].reshape(d, d), ref["G"]) < tol
    # normalised statistics are insensitive to the collapse regime: 1e-9
    assert relmax(m[2 + 2 * d:] / m[0], (ref["G"] / ref["S0"]).ravel()) < F64_TOL
    assert abs(out["result"]["eps"] - ref["eps"]) <= F64_TOL


# ==================================================================================================
# deep path (FP32)
# ==================================================================================================
def test_update_sample_weights_golden(dev):
    from rlvi_b200 import deep
    g = load_golden("deep_estep_n45000")
    res = cu(g["residuals"], dev)
    w = torch.ones(res.numel(), dtype=torch.float32, device=dev)
    deep.update_sample_weights(res, w)
    assert relmax(res.cpu().numpy(), g["residuals_after"]) < F32_TOL
    assert np.max(np.abs(w.cpu().numpy() - g["weights"])) < F32_TOL
    # epoch 2 starts from the truncated weights of epoch 1 (first-pass error against INCOMING weights)
    thr = deep.false_negative_criterion(w)
    assert thr.dim() == 0 and abs(float(thr) - float(g["threshold"])) <= F32_TOL
    w[w < thr] = 0
    mask_ref = g["weights_truncated"] > 0
    assert np.array_equal(w.cpu().numpy() > 0, mask_ref)        # identical selection mask
    res2 = cu(g["residuals2"], dev)
    deep.update_sample_weights(res2, w)
    assert np.max(np.abs(w.cpu().numpy() - g["weights2"])) < F32_TOL
    thr2 = deep.false_negative_criterion(w)
    assert abs(float(thr2) - float(g["threshold2"])) <= F32_TOL


def test_threshold_wrap_and_truncate(dev):
    from rlvi_b200 import deep, ops
    g = load_golden("deep_threshold_wrap")
    w = cu(g["weights"], dev)
    assert float(deep.false_negative_criterion(w)) == float(g["threshold"])     # quirk Q9
    rng = np.random.default_rng(4)
    for n in (1, 2, 17, 1000, 45000, 300000):
        wn = rng.beta(0.5, 0.5, size=n).astype(np.float32)
        wn[rng.integers(0, n)] = 1.0
        ref = deep_ref.false_negative_criterion(torch.from_numpy(wn))
        wt = cu(wn, dev)
        got = deep.false_negative_criterion(wt)
        assert float(got) == float(ref), n
        out = ops.fn_threshold(wt, prev_threshold=0.5, truncate=True)
        thr = max(0.5, float(ref))
        assert float(out) == np.float32(thr)
        exp = wn.copy()
        exp[exp < np.float32(thr)] = 0
        assert np.array_equal(wt.cpu().numpy(), exp)


def test_weighted_ce_golden_and_autograd(dev):
    from rlvi_b200 import deep
    g = load_golden("deep_wce_b512_c100")
    b = g["logits"].shape[0]
    n_train = 4 * b
    rng = np.random.default_rng(0)
    idx = rng.permutation(n_train)[:b].astype(np.int64)
    weights = np.zeros(n_train, dtype=np.float32)
    weights[idx] = g["batch_weights"]
    logits = cu(g["logits"], dev).requires_grad_(True)
    residuals = torch.zeros(n_train, dtype=torch.float32, device=dev)
    loss, correct = deep.weighted_cross_entropy(logits, cu(g["labels"], dev), cu(idx, dev), cu(weights, dev),
                                                residuals)
    (loss * 1.0).backward()
    assert abs(float(loss) - float(g["loss"])) < F32_TOL * abs(float(g["loss"]))
    assert relmax(residuals.cpu().numpy()[idx], g["per_sample"]) < F32_TOL
    assert not residuals.requires_grad                          # detached (quirk Q8)
    assert np.max(np.abs(logits.grad.cpu().numpy() - g["dlogits"])) < F32_TOL * np.max(np.abs(g["dlogits"]))
    # accuracy counts against torch.topk on the same logits
    lt = torch.from_numpy(g["logits"])
    top5 = lt.topk(5, dim=1).indices
    lab = torch.from_numpy(g["labels"])
    assert int(correct[0]) == int((top5[:, 0] == lab).sum())
    assert int(correct[1]) == int((top5 == lab[:, None]).any(dim=1).sum())


@pytest.mark.parametrize("b,c", [(1, 2), (7, 10), (100, 33), (513, 100), (64, 257), (32, 1000)])
def test_weighted_ce_shapes(dev, b, c):
    from rlvi_b200 import ops
    rng = np.random.default_rng(b * c)
    logits = torch.from_numpy((3 * rng.normal(size=(b, c))).astype(np.float32))
    labels = torch.from_numpy(rng.integers(0, c, size=b).astype(np.int64))
    w = torch.from_numpy(rng.random(b).astype(np.float32))
    per, loss, dl = deep_ref.weighted_ce(logits, labels, w)
    res = torch.zeros(b, dtype=torch.float32, device=dev)
    r = ops.wce_fwd_bwd(logits.to(dev), labels.to(dev), w.to(dev), res, want_per_sample=True)
    assert relmax(r["per_sample"].cpu().numpy(), per.numpy()) < F32_TOL
    assert torch.equal(res, r["per_sample"])
    assert abs(float(r["loss"]) - float(loss)) < F32_TOL * max(abs(float(loss)), 1e-3)
    assert np.max(np.abs(r["dlogits"].cpu().numpy() - dl.numpy())) < F32_TOL * max(float(dl.abs().max()), 1e-6)


def test_train_rlvi_epoch_matches_reference_restated(dev):
    """One epoch of the drop-in `train_rlvi` against the same epoch written with the oracle's torch ops
    on the same device (model + SGD identical): residuals, weights, accuracy and threshold agree."""
    from rlvi_b200 import deep
    torch.manual_seed(0)
    n_train, c, bs = 1024, 10, 128
    Xs = torch.randn(n_train, 20)
    ys = torch.randint(0, c, (n_train,))
    loader = [(Xs[i:i + bs], ys[i:i + bs], torch.arange(i, min(i + bs, n_train))) for i in range(0, n_train, bs)]

    def make():
        torch.manual_seed(1)
        m = torch.nn.Sequential(torch.nn.Linear(20, 32), torch.nn.ReLU(), torch.nn.Linear(32, c)).to(dev)
        return m, torch.optim.SGD(m.parameters(), lr=0.05, momentum=0.9)

    m1, o1 = make()
    res1 = torch.zeros(n_train, device=dev)
    w1 = torch.rand(n_train, device=dev) * 0.5 + 0.5
    w2, res2 = w1.clone(), res1.clone()
    acc1, thr1 = deep.train_rlvi(loader, m1, o1, res1, w1, True, 0)

    m2, o2 = make()
    correct = 0.0
    for xb, yb, ib in loader:
        xb, yb, ib = xb.to(dev), yb.to(dev), ib.to(dev)
        logits = m2(xb)
        correct += float((logits.argmax(1) == yb).float().sum() * (100.0 / yb.numel()))
        per = torch.nn.functional.cross_entropy(logits, yb, reduction="none")
        res2[ib] = per.detach()
        loss = (per * w2[ib]).mean()
        o2.zero_grad()
        loss.backward()
        o2.step()
    thr2 = deep_ref.epoch_tail(res2, w2, True, 0)
    assert abs(acc1 - correct / len(loader)) < 1e-9
    assert torch.allclose(res1, res2, rtol=1e-4, atol=1e-5)
    assert torch.allclose(w1, w2, rtol=1e-4, atol=1e-5)
    assert abs(float(thr1) - float(thr2)) < 1e-4
    assert torch.equal(w1 > 0, w2 > 0)


def test_train_rlvi_cuda_graph_matches_eager(dev):
    """deep.train_rlvi(cuda_graph={}) replays the per-batch body (model forward, fused weighted CE, backward, optimizer step;
    train_rlvi.py:84-97) as ONE captured CUDA graph; two epochs agree with the eager drop-in on the same model, data and
    optimizer: residuals, weights, parameters, accuracy, threshold.  The last batch is ragged and runs eagerly."""
    from rlvi_b200 import deep
    torch.manual_seed(0)
    n_train, c, bs = 1100, 10, 128
    Xs = torch.randn(n_train, 20)
    ys = torch.randint(0, c, (n_train,))
    loader = [(Xs[i:i + bs], ys[i:i + bs], torch.arange(i, min(i + bs, n_train))) for i in range(0, n_train, bs)]

    def run(graph):
        torch.manual_seed(1)
        m = torch.nn.Sequential(torch.nn.Linear(20, 32), torch.nn.ReLU(), torch.nn.Linear(32, c)).to(dev)
        o = torch.optim.Adam(m.parameters(), lr=1e-2, capturable=True)
        res = torch.zeros(n_train, device=dev)
        w = torch.ones(n_train, device=dev)
        state = {} if graph else None
        thr, accs = 0, []
        for _ in range(2):
            acc, thr = deep.train_rlvi(loader, m, o, res, w, True, thr, cuda_graph=state)
            accs.append(acc)
        return m, res, w, accs, thr, state

    m1, res1, w1, acc1, thr1, _ = run(False)
    m2, res2, w2, acc2, thr2, state = run(True)
    assert state["step"].graph is not None
    assert acc1 == pytest.approx(acc2, abs=1e-9)
    assert torch.allclose(res1, res2, rtol=1e-5, atol=1e-6)
    assert torch.allclose(w1, w2, rtol=1e-5, atol=1e-6)
    assert abs(float(thr1) - float(thr2)) < 1e-6
    for p1, p2 in zip(m1.parameters(), m2.parameters()):
        assert torch.allclose(p1, p2, rtol=1e-5, atol=1e-6)
