"""GPU tests for the tcgen05 layers: the bf16 hi/lo split GEMM (fp32 accumulation in TMEM) against
an fp64 reference.  Tolerance: 2e-5 of the output range (three-term split drops ~2^-18 relative
terms); the north-star bound for fp32 layer outputs is 1e-4."""
import pytest
import torch

from kd_pointcloud_b200 import functional as KF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
K = torch.ops.kdpc


def _ref(x, w, b=None, scale=None, shift=None, slope=1.0, clamp=None, residual=None):
    y = x.double() @ w.double().t()
    if scale is not None:
        y = y * scale.double()
    if shift is not None:
        y = y + shift.double()
    y = torch.where(y > 0, y, y * slope)
    if clamp is not None:
        y = y.clamp(*clamp)
    if residual is not None:
        y = y + residual.double()
    return y


def _err(a, b):
    return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


@pytest.mark.parametrize("m,n,k", [
    (128, 64, 64), (128, 128, 128), (1000, 64, 64), (300, 32, 32), (4096, 128, 2096), (777, 96, 320), (129, 16, 16),
    (65536, 128, 512), (5000, 256, 256), (200, 24, 100), (64, 256, 4144), (20000, 48, 80), (1, 64, 64),
    (3000, 512, 256), (1000, 2096, 128), (130, 8240, 256), (4096, 300, 64), (20000, 1072, 64),      # wide: one launch over column blocks
])
def test_linear_tc_matches_fp64(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).to(DEV)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(DEV)
    shift = torch.randn(n, generator=g).to(DEV)
    wp = K.pack_weight(w, 0, 0, 0)
    y = K.linear_tc(x, wp, n, None, shift, 0.1, 1.0, 0.0, None)
    assert y.shape == (m, n)
    assert _err(y, _ref(x, w, shift=shift, slope=0.1)) < 2e-5
    # run-to-run determinism
    assert torch.equal(y, K.linear_tc(x, wp, n, None, shift, 0.1, 1.0, 0.0, None))


def test_linear_tc_epilogue_variants_and_large_magnitudes():
    g = torch.Generator().manual_seed(0)
    m, n, k = 3000, 64, 192
    x = (torch.randn(m, k, generator=g) * 30).to(DEV)         # coordinates-scale activations
    w = (torch.randn(n, k, generator=g) * 0.2).to(DEV)
    scale, shift = torch.rand(n, generator=g).to(DEV) + 0.5, torch.randn(n, generator=g).to(DEV)
    res = torch.randn(m, n, generator=g).to(DEV)
    wp = K.pack_weight(w, 0, 0, 0)
    y = K.linear_tc(x, wp, n, scale, shift, 1.0, -150.0, 150.0, res)
    assert _err(y, _ref(x, w, scale=scale, shift=shift, clamp=(-150, 150), residual=res)) < 2e-5
    y = K.linear_tc(x, wp, n, None, None, 0.0, 1.0, 0.0, None)          # plain ReLU, no affine
    assert _err(y, _ref(x, w, slope=0.0)) < 2e-5


@pytest.mark.parametrize("m,n,k", [(5000, 32, 3), (4096, 3, 64), (100, 5, 7), (1000, 64, 10), (70001, 32, 3), (16384, 256, 3),
                                   (50000, 36, 2), (65539, 3, 64), (20000, 3, 128), (16385, 2, 36), (30000, 3, 32)])
def test_linear_simt_matches_fp64(m, n, k):
    """Generic CUDA-core kernel, and (m >= 16384) its two specialisations: k <= 4 with the weights in registers,
    n <= 4 with eight lanes per row."""
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).to(DEV)
    w = torch.randn(n, k, generator=g).to(DEV)
    shift = torch.randn(n, generator=g).to(DEV)
    scale = torch.rand(n, generator=g).to(DEV) + 0.5
    res = torch.randn(m, n, generator=g).to(DEV)
    y = K.linear_simt(x, w, None, shift, 1.0, -2.0, 2.0, res)
    assert _err(y, _ref(x, w, shift=shift, clamp=(-2, 2), residual=res)) < 1e-6
    y = K.linear_simt(x, w, scale, shift, 0.1, 1.0, 0.0, None)
    assert _err(y, _ref(x, w, scale=scale, shift=shift, slope=0.1)) < 1e-6
    assert torch.equal(y, K.linear_simt(x, w, scale, shift, 0.1, 1.0, 0.0, None))


def test_fused_linear_frontend_matches_torch_modules():
    torch.manual_seed(0)
    lin = torch.nn.Linear(320, 512).to(DEV)
    bn = torch.nn.BatchNorm1d(512).to(DEV).eval()
    bn.running_mean.normal_()
    bn.running_var.uniform_(0.5, 2.0)
    bn.weight.data.normal_(1, 0.1)
    bn.bias.data.normal_()
    x = torch.randn(4, 700, 320, device=DEV)
    with torch.no_grad():
        ref = torch.nn.functional.leaky_relu(bn(lin(x).reshape(-1, 512)).view(4, 700, 512).double(), 0.1)
        y = KF.fused_linear(x, lin.weight, lin.bias, bn, 0.1)           # N = 512: two column blocks
    assert y.shape == (4, 700, 512) and _err(y, ref) < 2e-5
    conv = torch.nn.Conv1d(64, 3, 1).to(DEV)
    res = torch.randn(4, 700, 3, device=DEV)
    xin = torch.randn(4, 700, 64, device=DEV)
    with torch.no_grad():
        ref = conv(xin.permute(0, 2, 1)).permute(0, 2, 1).double().clamp(-200, 200) + res.double()
        y = KF.fused_linear(xin, conv.weight, conv.bias, None, 1.0, (-200, 200), res)
    assert _err(y, ref) < 1e-5
    # weights are re-packed when they change in place
    with torch.no_grad():
        lin.weight.mul_(2.0)
        y2 = KF.fused_linear(x, lin.weight, lin.bias, None, 1.0)
        assert _err(y2, lin(x).double()) < 2e-5


# ------------------------------------------------------------------ fused PointConv / cost volume
def _gather(points, idx):
    B = points.shape[0]
    return points[torch.arange(B, device=points.device).view(B, 1, 1), idx.long()]


def _pointconv_ref64(cand, query, feats, idx, wn_convs, lin_w, scale, shift, slope):
    rel = _gather(cand, idx).double() - query.double().unsqueeze(2)                   # [B,S,K,3]
    grouped = torch.cat([rel, _gather(feats, idx).double()], dim=-1)                  # [B,S,K,3+D]
    w = rel
    for c in wn_convs:
        w = torch.relu(w @ c.weight.double().reshape(c.out_channels, -1).t() + c.bias.double())
    agg = torch.einsum("bskc,bskw->bscw", grouped, w).reshape(grouped.shape[0], grouped.shape[1], -1)
    y = agg @ lin_w.double().t()
    y = y * scale.double() + shift.double()
    return torch.where(y > 0, y, y * slope)


@pytest.mark.parametrize("B,N,S,D,Cout,KN", [(2, 1024, 1024, 32, 64, 9), (1, 700, 300, 128, 128, 9), (3, 512, 512, 320, 128, 9),
                                             (1, 8192, 8192, 128, 128, 9), (2, 8192, 2048, 64, 64, 16), (4, 256, 64, 512, 256, 16),
                                             (2, 512, 256, 256, 256, 16)])
def test_pointconv_fused_matches_fp64(B, N, S, D, Cout, KN):
    from kd_pointcloud_b200 import pointconv_util as P
    torch.manual_seed(B * 1000 + D)
    cand = (torch.rand(B, N, 3, device=DEV) * 4 - 2)
    query = cand[:, :S].contiguous() if S <= N else torch.rand(B, S, 3, device=DEV)
    feats = torch.randn(B, N, D, device=DEV)
    idx = K.knn(query, cand, KN)
    wn = P.WeightNet(3, 16).to(DEV)
    lin = torch.nn.Linear(16 * (D + 3), Cout).to(DEV)
    scale, shift = torch.rand(Cout, device=DEV) + 0.5, torch.randn(Cout, device=DEV)
    wp = K.pack_weight(lin.weight.detach(), 1, D, 16)
    params = KF._weightnet_host_params(wn.mlp_convs)
    with torch.no_grad():
        y = K.pointconv_fused(cand, query, feats, idx, params, wp, Cout, scale, shift, 0.1)
        ref = _pointconv_ref64(cand, query, feats, idx, wn.mlp_convs, lin.weight, scale, shift, 0.1)
    assert y.shape == (B, S, Cout)
    assert _err(y, ref) < 2e-5
    assert torch.equal(y, K.pointconv_fused(cand, query, feats, idx, params, wp, Cout, scale, shift, 0.1))
    # any per-cloud row order gives the same bits (rows are independent): a random permutation with a wider row
    # stride, and the Morton order the kNN leaves in the sort cache (what PointConv.forward passes)
    table = torch.stack([torch.randperm(S, device=DEV) for _ in range(B)]).int()
    wide = torch.cat([table, torch.full((B, 5), -1, dtype=torch.int32, device=DEV)], 1)
    assert torch.equal(y, K.pointconv_fused(cand, query, feats, idx, params, wp, Cout, scale, shift, 0.1, wide[:, :S]))
    if KF.ops.SORT_MIN_N <= N:
        KF._knn_compute(KN, cand, query)
        mo = KF.morton_order(query)
        assert mo is not None and torch.equal(mo.sort(dim=1).values, torch.arange(S, device=DEV).int().expand(B, S))
        assert torch.equal(y, K.pointconv_fused(cand, query, feats, idx, params, wp, Cout, scale, shift, 0.1, mo))


@pytest.mark.parametrize("B,N1,N2,D", [(2, 1024, 1024, 32), (1, 333, 500, 64), (2, 256, 256, 256), (1, 8192, 8192, 32),
                                        (1, 2048, 2048, 128), (3, 333, 500, 32), (5, 3, 40, 64), (2, 130, 64, 128),
                                        (3, 1367, 1500, 32), (2, 4099, 4100, 32)])
def test_costvol_fused_matches_fp64(B, N1, N2, D):
    torch.manual_seed(N1 + D)
    xyz1 = torch.rand(B, N1, 3, device=DEV) * 4
    xyz2 = torch.rand(B, N2, 3, device=DEV) * 4
    p1, p2 = torch.randn(B, N1, D, device=DEV), torch.randn(B, N2, D, device=DEV)
    idx = K.knn(xyz1, xyz2, 32)
    pos_w, pos_b = torch.randn(D, 3, device=DEV), torch.randn(D, device=DEV)
    w, b = torch.randn(D, D, device=DEV) / D ** 0.5, torch.randn(D, device=DEV)
    wp = K.pack_weight(w, 0, 0, 0)
    y = K.costvol_fused(xyz1, xyz2, p1, p2, idx, pos_w, pos_b, 0.1, wp, D, b, 0.1)
    rel = _gather(xyz2, idx).double() - xyz1.double().unsqueeze(2)
    x = _gather(p2, idx).double() + p1.double().unsqueeze(2) + rel @ pos_w.double().t() + pos_b.double()
    x = torch.where(x > 0, x, x * 0.1)
    z = x @ w.double().t() + b.double()
    z = torch.where(z > 0, z, z * 0.1)
    ref = z.max(dim=2)[0]
    assert y.shape == (B, N1, D)
    assert _err(y, ref) < 2e-5
    assert torch.equal(y, K.costvol_fused(xyz1, xyz2, p1, p2, idx, pos_w, pos_b, 0.1, wp, D, b, 0.1))
    # two row tiles per pipeline iteration against diag(W, W) (D = 32, >= 4096 points) == one tile per iteration, bit for bit
    from kd_pointcloud_b200 import _lib
    _lib.lib().kdpc_costvol_set_pairing(0)
    try:
        single = K.costvol_fused(xyz1, xyz2, p1, p2, idx, pos_w, pos_b, 0.1, wp, D, b, 0.1)
    finally:
        _lib.lib().kdpc_costvol_set_pairing(1)
    assert torch.equal(y, single)


def test_fused_layers_match_unfused_modules():
    """CrossLayerLight and the flow estimator through the fused kernels vs the same modules with the fused
    paths disabled (the kdpc op chain that the golden-vector tests pin against the reference)."""
    from kd_pointcloud_b200 import pointconv_util as P
    from kd_pointcloud_b200.synth import make_pairs, synthetic_state_dict
    d = make_pairs(2, 2048, seed=3, device=DEV)
    torch.manual_seed(1)
    cross = P.CrossLayerLight(32, 96, [64, 64], [64, 64]).to(DEV).eval()
    est = P.SceneFlowEstimatorResidual(128, 64).to(DEV).eval()
    for m in (cross, est):
        m.load_state_dict(synthetic_state_dict(m.state_dict(), 11))
    f1, f2 = torch.randn(2, 2048, 96, device=DEV), torch.randn(2, 2048, 96, device=DEV)
    feats, cost = torch.randn(2, 2048, 128, device=DEV), torch.randn(2, 2048, 64, device=DEV)
    with torch.no_grad():
        fused = cross.forward_pm(d["pos1"], d["pos2"], f1, f2) + est.forward_pm(d["pos1"], feats, cost)
        saved = KF.FUSED_POINTCONV_K, P.FUSED_COSTVOL
        KF.FUSED_POINTCONV_K, P.FUSED_COSTVOL = (), False
        try:
            plain = cross.forward_pm(d["pos1"], d["pos2"], f1, f2) + est.forward_pm(d["pos1"], feats, cost)
        finally:
            KF.FUSED_POINTCONV_K, P.FUSED_COSTVOL = saved
    for a, b in zip(fused, plain):
        assert _err(a, b.double()) < 5e-5


@pytest.mark.parametrize("M,Kin,N", [(4096, 128, 128), (1000, 2096, 128), (300, 64, 512), (8192, 3120, 128), (130, 16, 16)])
def test_linear_tc_autograd_matches_fp64(M, Kin, N):
    """Training path: forward and input gradient on tcgen05 (bf16 hi/lo), dW / db by torch; against an fp64 reference."""
    torch.manual_seed(M + N)
    x = torch.randn(2, M // 2, Kin, device=DEV, requires_grad=True)
    lin = torch.nn.Linear(Kin, N).to(DEV)
    gy = torch.randn(2, M // 2, N, device=DEV)
    assert KF.linear_tc_autograd_available(x, lin.weight)
    y = KF.linear_tc_autograd(x, lin.weight, lin.bias)
    y.backward(gy)
    xd = x.detach().double().requires_grad_(True)
    wd, bd = lin.weight.detach().double().requires_grad_(True), lin.bias.detach().double().requires_grad_(True)
    yd = torch.nn.functional.linear(xd, wd, bd)
    yd.backward(gy.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    assert rel(y, yd) < 1e-5 and rel(x.grad, xd.grad) < 1e-5
    assert rel(lin.weight.grad, wd.grad) < 1e-5 and rel(lin.bias.grad, bd.grad) < 1e-5


def test_wide_linear_blocks_written_in_place():
    """N > 256: column blocks go straight into the output (no torch.cat), same values as the per-block results."""
    torch.manual_seed(5)
    x = torch.randn(3, 700, 256, device=DEV)
    lin = torch.nn.Linear(256, 512).to(DEV)
    with torch.no_grad():
        y = KF.fused_linear(x, lin.weight, lin.bias, None, 0.1)
        ref = torch.nn.functional.leaky_relu(torch.nn.functional.linear(x.double(), lin.weight.double(), lin.bias.double()), 0.1)
        lo = KF.fused_linear(x, lin.weight[:256], lin.bias[:256], None, 0.1)
    assert ((y.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5
    assert torch.equal(y[..., :256], lo)


@pytest.mark.parametrize("rows,k,n,ctot,c0,otot,o0", [
    ((16, 8192), 32, 64, 64, 0, 64, 0),          # level0_2 reading f_l0 out of the level-0 concatenation buffer
    ((16, 2048), 64, 128, 96, 0, 128, 0),        # level1_1 (row stride 96)
    ((16, 256), 256, 512, 320, 0, 512, 0),       # level3_1: strided operand of a WIDE layer (column blocks)
    ((16, 512), 128, 64, 128, 0, 192, 128),      # deconv3_2 writing the second column block
    ((4, 300), 64, 32, 100, 36, 52, 20),         # ragged rows, both views offset (16-byte aligned)
    ((2, 4096), 256, 256, 256, 0, 512, 256),     # small-M layer on the split-N plan, strided result
    ((1, 64), 4144, 256, 4144, 0, 320, 0),       # split-K plan (workspace + reduce) into a strided result
])
def test_linear_reads_and_writes_column_blocks(rows, k, n, ctot, c0, otot, o0):
    """fused_linear on row-strided views (operand = column block of a wider tensor, result = column block of another)
    is bit-identical to the contiguous call, and leaves the rest of the destination untouched."""
    g = torch.Generator().manual_seed(k + n + ctot)
    big = torch.randn(*rows, ctot, generator=g).to(DEV)
    lin = torch.nn.Linear(k, n).to(DEV)
    xv = big[..., c0:c0 + k]
    with torch.no_grad():
        ref = KF.fused_linear(xv.contiguous(), lin.weight, lin.bias, None, 0.1)
        dest = torch.full((*rows, otot), 7.0, device=DEV)
        got = KF.fused_linear(xv, lin.weight, lin.bias, None, 0.1, out=dest[..., o0:o0 + n])
        only_x = KF.fused_linear(xv, lin.weight, lin.bias, None, 0.1)
    assert got.data_ptr() == dest[..., o0:o0 + n].data_ptr()
    assert torch.equal(dest[..., o0:o0 + n], ref) and torch.equal(only_x, ref)
    keep = torch.ones(otot, dtype=torch.bool)
    keep[o0:o0 + n] = False
    assert bool((dest[..., keep.to(DEV)] == 7.0).all())
    # batch halves of one tensor: the second half starts mid-buffer
    with torch.no_grad():
        both = torch.empty(2 * rows[0], *rows[1:], n, device=DEV)
        KF.fused_linear(xv, lin.weight, lin.bias, None, 0.1, out=both[rows[0]:])
        KF.fused_linear(xv, lin.weight, lin.bias, None, 0.1, out=both[:rows[0]])
    assert torch.equal(both[:rows[0]], ref) and torch.equal(both[rows[0]:], ref)
    assert KF.joined(both[:rows[0]], both[rows[0]:]) is both


@pytest.mark.parametrize("m,n,k", [(65536, 128, 128), (4096, 64, 2096), (1000, 32, 35), (128 * 3 + 5, 256, 256), (50, 16, 16),
                                   (8192, 128, 1072), (300, 200, 520), (131072, 32, 32), (65536, 32, 3), (32768, 64, 63),
                                   (16384, 64, 64), (6400, 48, 20), (64, 8, 3), (70001, 128, 3), (4099, 256, 3), (1000, 36, 2)])
def test_linear_dw_matches_fp64(m, n, k):
    """dW = dY^T X on tcgen05 with MN-major operand tiles (csrc/dw_tc.cu) against an fp64 product: ragged row counts,
    N < 128 (zero-padded tile rows), K tails that are not multiples of 16 / 64 / 256, several split counts."""
    g = torch.Generator().manual_seed(m + n + k)
    dy = torch.randn(m, n, generator=g).to(DEV)
    x = (torch.randn(m, k, generator=g) + 0.5).to(DEV)
    dw, db = torch.ops.kdpc.linear_dw(dy, x, True)
    ref = dy.double().t() @ x.double()
    err = ((dw.double() - ref).abs().max() / ref.abs().max()).item()
    assert dw.shape == (n, k) and err < 2e-5, err
    ref_b = dy.double().sum(0)
    assert db.shape == (n,) and ((db.double() - ref_b).abs().max() / ref_b.abs().max()).item() < 2e-5
    dw2, db2 = torch.ops.kdpc.linear_dw(dy, x, True)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)                     # deterministic
    dw3, db3 = torch.ops.kdpc.linear_dw(dy, x, False)                        # without the bias column: same weight gradient
    assert db3.numel() == 0 and ((dw3.double() - ref).abs().max() / ref.abs().max()).item() < 2e-5
    # narrow contiguous shapes take the bulk-TMA row fetch: same bits as the register-staged fetch
    from kd_pointcloud_b200 import _lib
    _lib.lib().kdpc_linear_dw_set_async(0)
    try:
        dw4, db4 = torch.ops.kdpc.linear_dw(dy, x, True)
    finally:
        _lib.lib().kdpc_linear_dw_set_async(1)
    if k + 1 > 4:
        assert torch.equal(dw, dw4) and torch.equal(db, db4)
    else:       # tiny input width: fp32 column sums on the CUDA cores instead of the bf16 hi/lo product - both inside the budget
        assert ((dw4.double() - ref).abs().max() / ref.abs().max()).item() < 2e-5
        assert ((db4.double() - ref_b).abs().max() / ref_b.abs().max()).item() < 2e-5


def test_linear_tc_autograd_uses_the_tcgen05_weight_gradient():
    from kd_pointcloud_b200 import functional as KF
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 700, 96, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(48, 96, generator=g) * 0.1).to(DEV).requires_grad_(True)
    b = torch.randn(48, generator=g).to(DEV).requires_grad_(True)
    go = torch.randn(2, 700, 48, generator=g).to(DEV)
    outs = {}
    for tc_dw in (True, False):
        KF.USE_TC_DW = tc_dw
        try:
            for t in (x, w, b):
                t.grad = None
            KF.linear_tc_autograd(x, w, b).backward(go)
            outs[tc_dw] = (x.grad.clone(), w.grad.clone(), b.grad.clone())
        finally:
            KF.USE_TC_DW = True
    ref_w = go.double().reshape(-1, 48).t() @ x.detach().double().reshape(-1, 96)
    for tc_dw in (True, False):
        assert ((outs[tc_dw][1].double() - ref_w).abs().max() / ref_w.abs().max()).item() < 2e-5
    assert torch.equal(outs[True][0], outs[False][0])
    ref_b = go.double().reshape(-1, 48).sum(0)
    for tc_dw in (True, False):
        assert ((outs[tc_dw][2].double() - ref_b).abs().max() / ref_b.abs().max()).item() < 2e-5


def test_small_channel_linear_autograd_matches_torch():
    """3 -> 32 (cost-volume positional encoding) and 64 -> 3 (flow head) layers of the training path: SIMT forward / dX,
    tcgen05 dW / db (functional._LinearSmall) against torch autograd."""
    from kd_pointcloud_b200 import functional as KF
    g = torch.Generator().manual_seed(5)
    for k, n, lead in ((3, 32, (2, 300, 32)), (64, 3, (2, 4099)), (3, 64, (1, 5000, 16))):
        x0 = torch.randn(*lead, k, generator=g).to(DEV)
        w0 = (torch.randn(n, k, generator=g) * 0.3).to(DEV)
        b0 = torch.randn(n, generator=g).to(DEV)
        go = torch.randn(*lead, n, generator=g).to(DEV)
        res = []
        for small in (True, False):
            x, w, b = (t.clone().requires_grad_(True) for t in (x0, w0, b0))
            assert KF.linear_small_autograd_available(x, w)
            y = KF.linear_small_autograd(x, w, b) if small else torch.nn.functional.linear(x, w, b)
            y.backward(go)
            res.append((y.detach(), x.grad, w.grad, b.grad))
        for a, r in zip(res[0], res[1]):
            assert a.shape == r.shape and ((a - r).abs().max() / r.abs().max()).item() < 2e-5


@pytest.mark.parametrize("B,N1,N2,D", [(2, 700, 650, 32), (1, 4096, 4096, 32), (2, 515, 700, 64), (1, 2048, 2048, 64)])
def test_costvol_autograd_matches_fp64(B, N1, N2, D):
    """Training path of the 8192- and 2048-point cost volumes (K = 32, D = D' = 32 / 64): fused tcgen05 forward + the recomputing arg-max
    backward (csrc/costvol_grad.cu) against fp64 autograd of the reference's op chain (pointconv_util.py:1826-1850);
    every input and parameter gradient, run-to-run identical."""
    from kd_pointcloud_b200 import functional as KF
    torch.manual_seed(N1)
    xyz1 = (torch.rand(B, N1, 3, device=DEV) * 4).requires_grad_(True)
    xyz2 = (torch.rand(B, N2, 3, device=DEV) * 4).requires_grad_(True)
    p1 = torch.randn(B, N1, D, device=DEV, requires_grad=True)
    p2 = torch.randn(B, N2, D, device=DEV, requires_grad=True)
    idx = K.knn(xyz1.detach(), xyz2.detach(), 32)
    pos = torch.nn.Conv2d(3, D, 1).to(DEV)
    conv = torch.nn.Conv2d(D, D, 1).to(DEV)
    go = torch.randn(B, N1, D, device=DEV)
    params = [xyz1, xyz2, p1, p2, pos.weight, pos.bias, conv.weight, conv.bias]

    def run():
        for t in params:
            t.grad = None
        assert KF.costvol_autograd_available(p1, idx, conv, 0.1)
        out = KF.costvol_autograd(xyz1, xyz2, p1, p2, idx, pos, 0.1, conv, 0.1)
        out.backward(go)
        return out.detach().clone(), [t.grad.detach().clone() for t in params]

    out, grads = run()
    dd = [t.detach().double().requires_grad_(True) for t in params]
    x1, x2, q1, q2, pw, pb, cw, cb = dd
    gi = idx.long()
    bi = torch.arange(B, device=DEV).view(B, 1, 1)
    rel = x2[bi, gi] - x1.unsqueeze(2)
    x = q2[bi, gi] + q1.unsqueeze(2) + rel @ pw.reshape(D, 3).t() + pb
    h = torch.where(x > 0, x, x * 0.1)
    z = h @ cw.reshape(D, D).t() + cb
    z = torch.where(z > 0, z, z * 0.1)
    ref = z.max(dim=2)[0]
    ref.backward(go.double())
    assert _err(out, ref.detach()) < 2e-5
    for name, g, r in zip(("xyz1", "xyz2", "p1", "p2", "pos.w", "pos.b", "conv.w", "conv.b"), grads, dd):
        rg = r.grad.reshape(g.shape)
        l2 = ((g.double() - rg).norm() / rg.norm().clamp_min(1e-30)).item()
        assert l2 < 1e-4, (name, l2)
    out2, grads2 = run()
    assert torch.equal(out, out2) and all(torch.equal(a, b) for a, b in zip(grads, grads2))
