"""GPU: the hand-written tcgen05 dense kernels of the PPO update (csrc/tg_gemm.cu, include/fortattack_train.h) against
float64 products, and the optimizer kernel against torch's clip_grad_norm_ + Adam (rlcore/algo/ppo.py:189-192).
Gate: fp32-grade -- the split-operand products must be as close to the float64 result as an fp32 GEMM is."""
from importlib import import_module

import pytest
import torch

pytestmark = pytest.mark.gpu
PKG = "emergent-multiagent-strategies_b200"
fused = import_module(PKG + ".rlcore.fused")
_capi = import_module(PKG + "._capi")


def _rows(gen, rows, cols, wide):
    x = torch.randn(rows, cols, generator=gen)
    if wide:            # rows of very different magnitude (gradient rows of dead / alive agents, raw headings next to flags)
        x = x * torch.pow(10.0, torch.rand(rows, 1, generator=gen) * 9 - 6)
        x[::7] = 0.0
    return x


@pytest.mark.parametrize("rows,K,N", [(1000, 128, 128), (777, 256, 128), (300, 128, 256), (5000, 64, 64), (129, 64, 128),
                                      (1000, 6, 64), (500, 128, 8), (500, 128, 1), (500, 8, 128), (500, 1, 128),
                                      (1, 128, 128), (128, 128, 128), (40000, 128, 128), (20011, 256, 128)])
@pytest.mark.parametrize("wide", [False, True])
def test_linear_matches_float64(rows, K, N, wide):
    gen = torch.Generator().manual_seed(rows * 7 + K + N)
    x = _rows(gen, rows, K, wide).cuda()
    W = (torch.randn(N, K, generator=gen) / K ** 0.5).cuda()
    b = torch.randn(N, generator=gen).cuda()
    ref = x.double() @ W.double().t()
    bound = (x.double().abs() @ W.double().abs().t())                   # what an fp32 dot product's error is relative to
    y = fused.tg_linear(x, fused.tg_pack(W, False))
    fused.tg_check_status("cuda:0")
    err = ((y.double() - ref).abs() / (bound + 1e-30)).max()
    assert float(err) < 2e-6, float(err)
    # bias + ReLU epilogue, transposed pack (B given as [K, N]), accumulate into an existing output
    y2 = fused.tg_linear(x, fused.tg_pack(W.t().contiguous(), True), b, relu=True)
    assert float(((y2.double() - torch.relu(ref + b.double())).abs() / (bound + b.double().abs() + 1e-30)).max()) < 2e-6
    base = torch.randn(rows, N, generator=gen).cuda()
    y3 = fused.tg_linear(x, fused.tg_pack(W, False), out=base.clone(), accumulate=True)
    assert float(((y3.double() - (ref + base.double())).abs() / (bound + base.double().abs() + 1e-30)).max()) < 2e-6


@pytest.mark.parametrize("rows,K,N,ld", [(1000, 128, 128, 128), (129, 64, 128, 64), (5000, 64, 64, 192), (20011, 128, 64, 256),
                                         (31, 128, 8, 128), (40000, 128, 128, 128), (4097, 128, 1, 132), (3000, 256, 128, 256),
                                         (2777, 128, 256, 128), (1500, 6, 64, 8), (999, 128, 20, 128)])
def test_linear_staged_form_equals_register_form(rows, K, N, ld):
    """The bulk-copy staged loader (default for K = 64 / 128) and the register loader convert the same values with the same
    arithmetic, and the TMA tensor-store epilogue writes what the STG epilogue writes: outputs are bit-equal across the four
    combinations, for contiguous and strided rows / outputs, ragged tails, partial column blocks, with bias / ReLU / accumulate."""
    L = _capi.lib()
    gen = torch.Generator().manual_seed(rows + K + N + ld)
    big = _rows(gen, rows, ld, True).cuda()
    x = big[:, ld - K:]
    W = (torch.randn(N, K, generator=gen) / K ** 0.5).cuda()
    b = torch.randn(N, generator=gen).cuda()
    base = torch.randn(rows, N, generator=gen).cuda()
    outs = []
    for staged, tma in ((1, 1), (1, 0), (0, 1), (0, 0)):          # loader form x epilogue form (TMA tensor stores / STG)
        L.tg_debug_staged(staged)
        L.tg_debug_tma_out(tma)
        try:
            pack = fused.tg_pack(W, False)
            wide = torch.full((rows, N + 4), 7.0, device="cuda")    # strided output: the column beside it must stay untouched
            fused.tg_linear(x, pack, b, relu=True, out=wide[:, :N])
            outs.append((fused.tg_linear(x, pack), fused.tg_linear(x, pack, b, relu=True),
                         fused.tg_linear(x, pack, out=base.clone(), accumulate=True), wide))
        finally:
            L.tg_debug_staged(1)
            L.tg_debug_tma_out(1)
    fused.tg_check_status("cuda:0")
    for o in outs[1:]:
        for a, c in zip(outs[0], o):
            assert torch.equal(a, c)
    assert torch.equal(outs[0][3][:, :N], outs[0][1]) and float((outs[0][3][:, N:] - 7.0).abs().max()) == 0.0
    ref = x.double() @ W.double().t()
    bound = x.double().abs() @ W.double().abs().t()
    assert float(((outs[0][0].double() - ref).abs() / (bound + 1e-30)).max()) < 2e-6


def test_linear_strided_operands():
    """x and out as column slices of wider row-major tensors (the [h | mixed] buffer of a message round)."""
    gen = torch.Generator().manual_seed(3)
    big = torch.randn(3000, 256, generator=gen).cuda()
    W = torch.randn(128, 128, generator=gen).cuda() / 11
    ref = big[:, 128:].double() @ W.double().t()
    out = torch.zeros(3000, 256, device="cuda")
    fused.tg_linear(big[:, 128:], fused.tg_pack(W, False), out=out[:, :128])
    assert float((out[:, :128].double() - ref).abs().max()) < 2e-5 and float(out[:, 128:].abs().max()) == 0.0
    Wv = torch.randn(128, 256, generator=gen).cuda()                       # weight given as a strided view: update.0.weight[:, 128:]
    y = fused.tg_linear(big[:, :128], fused.tg_pack(Wv[:, 128:], False))
    assert float((y.double() - big[:, :128].double() @ Wv[:, 128:].double().t()).abs().max()) < 1e-4


def _wgrad_err(x, y):
    ref = x.double().t() @ y.double()
    bound = x.double().abs().t() @ y.double().abs()
    got = fused.tg_wgrad(x, y)
    return float(((got.double() - ref).abs() / (bound + 1e-30)).max()), float((got.double() - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("rows,a,b", [(5000, 256, 128), (4099, 128, 256), (1000, 128, 128), (33, 64, 8), (70000, 256, 64)])
def test_wgrad_block_heights(rows, a, b):
    """Both block heights of tg_wgrad (64-row blocks, and 32-row blocks = the default when an operand is wider than 128 columns)
    against float64, ragged row counts included."""
    L = _capi.lib()
    gen = torch.Generator().manual_seed(rows + a + b)
    x, y = _rows(gen, rows, a, False).cuda(), _rows(gen, rows, b, True).cuda()
    for wr in (32, 64, 0):
        L.tg_debug_wgrad_rows(wr)
        try:
            e_rel, e_abs = _wgrad_err(x, y)
        finally:
            L.tg_debug_wgrad_rows(0)
        fused.tg_check_status("cuda:0")
        assert e_rel < 2e-6, (wr, e_rel, e_abs)


@pytest.mark.parametrize("rows,a,b,lda,ldb", [(5000, 128, 128, 128, 128), (4099, 256, 128, 256, 128), (3001, 128, 256, 128, 256),
                                               (1000, 64, 64, 128, 64), (33, 64, 8, 64, 8), (70000, 128, 64, 256, 192), (777, 8, 128, 8, 128)])
def test_wgrad_staged_form_equals_register_form(rows, a, b, lda, ldb):
    """The tensor-map staged loaders of tg_wgrad and the register loaders split the same values into the same operand blocks and
    partition the rows the same way: results are bit-equal, for contiguous and strided operands and ragged row counts."""
    L = _capi.lib()
    gen = torch.Generator().manual_seed(rows + a + b)
    x, y = _rows(gen, rows, lda, True).cuda()[:, lda - a:], _rows(gen, rows, ldb, False).cuda()[:, :b]
    outs = []
    for on in (1, 0):
        L.tg_debug_wgrad_staged(on)
        try:
            outs.append(fused.tg_wgrad(x, y))
        finally:
            L.tg_debug_wgrad_staged(1)
    fused.tg_check_status("cuda:0")
    assert torch.equal(outs[0], outs[1])
    ref = x.double().t() @ y.double()
    bound = x.double().abs().t() @ y.double().abs()
    assert float(((outs[0].double() - ref).abs() / (bound + 1e-30)).max()) < 2e-6


def test_wgrad_descriptor_probe():
    """Prints the error of the built-in MN-major descriptor fields and of the swapped pair (diagnostic for the layout)."""
    gen = torch.Generator().manual_seed(0)
    x, y = torch.randn(640, 128, generator=gen).cuda(), torch.randn(640, 128, generator=gen).cuda()
    L = _capi.lib()
    res = {}
    for name, (lbo, sbo) in (("lbo=128,sbo=1040", (128, 1040)), ("lbo=1040,sbo=128", (1040, 128))):
        L.tg_debug_wgrad_desc(lbo, sbo)
        res[name] = _wgrad_err(x, y)
        torch.cuda.synchronize()
    L.tg_debug_wgrad_desc(0, 0)
    print("tg_wgrad descriptor probe (err/bound, err/max):", res)
    assert res["lbo=128,sbo=1040"][0] < 2e-6, res


@pytest.mark.parametrize("rows,a,b", [(5000, 128, 128), (3001, 256, 128), (1000, 128, 256), (2000, 64, 6), (2000, 8, 128),
                                      (2000, 1, 128), (64, 128, 128), (63, 64, 64), (100000, 128, 128), (196608, 256, 128)])
@pytest.mark.parametrize("wide", [False, True])
def test_wgrad_matches_float64(rows, a, b, wide):
    gen = torch.Generator().manual_seed(rows + a * 3 + b)
    x, y = _rows(gen, rows, a, wide).cuda(), _rows(gen, rows, b, False).cuda()
    e_bound, e_max = _wgrad_err(x, y)
    fused.tg_check_status("cuda:0")
    assert e_bound < 2e-6, (e_bound, e_max)
    # bit-reproducible (fixed-order partial sums, no atomics)
    assert torch.equal(fused.tg_wgrad(x, y), fused.tg_wgrad(x, y))


def test_wgrad_strided_operands():
    gen = torch.Generator().manual_seed(5)
    hm = torch.randn(4000, 256, generator=gen).cuda()
    dg = torch.randn(4000, 128, generator=gen).cuda()
    got = fused.tg_wgrad(hm[:, :128], dg)
    ref = hm[:, :128].double().t() @ dg.double()
    assert float((got.double() - ref).abs().max() / ref.abs().max()) < 1e-6


def test_adam_step_matches_torch():
    torch.manual_seed(0)
    shapes = [(128, 256), (128,), (64, 6), (1, 128, 128), (8, 128), (1,), (64, 128)]
    ps = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    ref = torch.optim.Adam(qs, lr=1e-3)
    own = fused.TgAdam(ps, lr=1e-3)
    for it in range(6):
        scale = 10.0 if it % 2 == 0 else 1e-3                     # clipped and unclipped steps
        for k, (p, q) in enumerate(zip(ps, qs)):
            g = torch.randn_like(p) * scale
            # tensor 5 never receives a gradient (like the reference's unused oppUpdate layer, mpnn.py:44-45): skipped
            p.grad, q.grad = (None, None) if k == 5 else (g.clone(), g.clone())
        tn = torch.nn.utils.clip_grad_norm_(qs, 0.5)
        ref.step()
        own.step(0.5)
        assert abs(float(own.total_norm) - float(tn)) < 1e-4 * float(tn)
        for p, q in zip(ps, qs):
            assert float((p - q).abs().max()) < 2e-6, it
    assert int(own.step_count) == 6


def test_dense_layers_match_library_path():
    """fused.linear / matmul / matmul_nt / message_round with DENSE='tcgen05' against the same functions on cuBLAS:
    outputs and every gradient."""
    gen = torch.Generator().manual_seed(1)
    n, B, d = 3, 700, 128
    h = torch.randn(n * B, d, generator=gen).cuda().requires_grad_()
    Mqk = (torch.randn(d, d, generator=gen) / d ** 0.5).cuda().requires_grad_()
    Wc = (torch.randn(2 * d, d, generator=gen) / (2 * d) ** 0.5).cuda().requires_grad_()
    bias = torch.randn(d, generator=gen).cuda().requires_grad_()
    W2 = (torch.randn(8, d, generator=gen) / d ** 0.5).cuda().requires_grad_()
    b2 = torch.zeros(8, device="cuda", requires_grad=True)
    gout = torch.randn(n * B, 8, generator=gen).cuda()
    res = {}
    rec = {"mode": "record", "outs": []}
    rep = {"mode": "replay", "outs": rec["outs"], "pos": 0, "flips": 0}
    for mode in ("cublas", "tcgen05"):
        fused.DENSE = mode
        fused.RELU_TRACE = rec if mode == "cublas" else rep      # ReLU decisions of the backward pinned to the library run's
        try:
            for t in (h, Mqk, Wc, bias, W2, b2):
                t.grad = None
            x, attn = fused.message_round(h, Mqk, Wc, bias, n, d ** -0.5)
            x = fused.matmul_nt(fused.matmul(x, Mqk), Wc[:d])
            y = fused.linear(x, W2, b2)
            (y * gout).sum().backward()
            res[mode] = [y.detach().clone()] + [t.grad.clone() for t in (h, Mqk, Wc, bias, W2, b2)]
        finally:
            fused.DENSE, fused.RELU_TRACE = "tcgen05", None
    fused.tg_check_status("cuda:0")
    assert rep["pos"] == len(rec["outs"]) == 1 and rep["flips"] <= 2, rep["flips"]
    for a, b in zip(res["cublas"], res["tcgen05"]):
        assert float((a - b).abs().max()) < 2e-4 * max(1.0, float(a.abs().max())), float((a - b).abs().max())
