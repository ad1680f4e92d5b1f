"""GPU parity tests, kernel level: every libkdpc kernel (called through torch.ops.kdpc -> C ABI)
against the CPU oracle on the same seeded inputs.  Index/integer work must be bit-exact; fp32
arithmetic that the reference pins (distances, three_interpolate) must be bit-exact too; the rest
is checked to 1e-5 relative (tolerance stated per test)."""
import numpy as np
import pytest
import torch

from oracle import layers_ref as O
from kd_pointcloud_b200 import functional as KF
from kd_pointcloud_b200.synth import make_pairs

pytestmark = pytest.mark.gpu
K = torch.ops.kdpc
DEV = "cuda:0"


def _cloud(b, n, seed, mode):
    if mode == "grid":                                   # integer grid: exact ties everywhere
        g = torch.Generator().manual_seed(seed)
        return torch.randint(0, 7, (b, n, 3), generator=g).float()
    if mode == "dup":
        return make_pairs(b, n, seed=seed, duplicates=0.05)["pos1"]
    if mode == "cm":
        return make_pairs(b, n, seed=seed, quantize=0.01)["pos1"]
    if mode == "kitti":
        return make_pairs(b, n, seed=seed, kind="kitti")["pos1"]
    return make_pairs(b, n, seed=seed)["pos1"]


# ------------------------------------------------------------------------------------ FPS
@pytest.mark.parametrize("n,m,mode", [
    (8192, 2048, "ft3d"), (8192, 2048, "dup"), (2048, 512, "cm"), (512, 256, "grid"), (256, 64, "kitti"),
    (1000, 100, "grid"), (3000, 64, "dup"), (5000, 33, "ft3d"), (37, 9, "grid"), (20, 20, "grid"), (10000, 50, "dup"),
    (64, 1, "ft3d"),
])
def test_fps_bit_exact(n, m, mode):
    xyz = _cloud(2, n, 100 + n, mode)
    got = K.fps(xyz.to(DEV), m).cpu()
    assert got.dtype == torch.int32 and got.shape == (2, m)
    assert torch.equal(got, O.furthest_point_sample(xyz, m))


def test_fps_cluster_kernel_equals_single_cta_kernel():
    """8192- and 2048-point clouds run on a cluster of 8 CTAs per cloud (DSMEM arg-max exchange): identical
    indices AND identical final min-distance field to the one-CTA kernel, on tie-heavy clouds too."""
    from kd_pointcloud_b200 import _lib
    L = _lib.lib()
    for n, m, mode in ((8192, 2048, "ft3d"), (8192, 2048, "grid"), (2048, 512, "dup"), (4096, 1000, "cm"), (16384, 300, "kitti")):
        xyz = _cloud(5, n, 900 + n, mode).to(DEV)
        a = K.fps(xyz, m)
        L.kdpc_fps_set_cluster(0)
        try:
            b = K.fps(xyz, m)
        finally:
            L.kdpc_fps_set_cluster(1)
        assert torch.equal(a, b), (n, m, mode)
        assert torch.equal(a[:2].cpu(), O.furthest_point_sample(xyz[:2].cpu(), m))


def test_fps_batch_of_16_clouds_and_determinism():
    xyz = _cloud(16, 2048, 7, "dup").to(DEV)
    a, b = K.fps(xyz, 512), K.fps(xyz, 512)
    assert torch.equal(a, b)
    assert torch.equal(a[5:7].cpu(), O.furthest_point_sample(xyz[5:7].cpu(), 512))


# ------------------------------------------------------------------------------------ kNN
@pytest.mark.parametrize("s,n,k,mode", [
    (8192, 8192, 32, "ft3d"), (2048, 8192, 16, "kitti"), (8192, 2048, 3, "ft3d"), (512, 512, 9, "grid"),
    (300, 1000, 16, "dup"), (64, 256, 16, "cm"), (100, 777, 1, "ft3d"), (100, 777, 5, "grid"),
    (100, 777, 10, "dup"), (50, 513, 7, "ft3d"), (50, 40, 20, "grid"), (33, 2000, 24, "kitti"), (4, 32, 32, "ft3d"),
])
def test_knn_bit_exact_indices_and_distances(s, n, k, mode):
    cand = _cloud(1, n, 200 + n, mode)
    if s == n:
        query = cand.clone()                              # self-query (PointConv, flow estimator)
    else:
        query = _cloud(1, max(s, 8), 300 + s, mode)[:, :s].contiguous()
    idx, dist = K.knn_dist(query.to(DEV), cand.to(DEV), k)
    ref_i, ref_d = O.knn_with_dist(k, cand, query)
    assert torch.equal(idx.cpu(), ref_i)                  # same order too: ascending (distance, index)
    assert torch.equal(dist.cpu(), ref_d)                 # fp32 bit-exact: same rounding sequence
    i32, i64 = K.knn64(query.to(DEV), cand.to(DEV), k)
    assert i64.dtype == torch.int64 and torch.equal(i64.cpu(), ref_i.long()) and torch.equal(i32, idx)


@pytest.mark.parametrize("b,s,n,k,mode", [
    (8, 8192, 8192, 32, "ft3d"), (4, 8192, 8192, 32, "dup"), (4, 8192, 8192, 9, "cm"), (2, 8192, 8192, 3, "grid"),
    (4, 2048, 8192, 16, "kitti"), (4, 8192, 2048, 3, "ft3d"), (3, 1000, 5000, 10, "dup"), (2, 16384, 16384, 16, "ft3d"),
    (2, 300, 256, 32, "grid"), (2, 257, 999, 5, "cm"), (8, 8192, 8192, 3, "dup"), (8, 8192, 8192, 1, "cm"), (8, 2048, 8192, 4, "grid"),
    (5, 1999, 3001, 3, "kitti"), (8, 8192, 2048, 2, "dup"),
])
def test_knn_pruned_search_equals_brute_force(b, s, n, k, mode):
    """Size-independent property at full size: the best-first tile-pruned search (sorted clouds, conservative
    bound) returns exactly what scanning every candidate returns - including clouds full of exact ties."""
    cand = _cloud(b, n, 400 + n, mode).to(DEV)
    query = cand if s == n else _cloud(b, s, 500 + s, mode).to(DEV)
    brute = K.knn_bruteforce(query, cand, k)
    assert torch.equal(K.knn(query, cand, k), brute)
    qs, cs = K.spatial_sort(query), K.spatial_sort(cand)
    assert torch.equal(K.knn_sorted(qs, cs, b, s, n, k), brute)
    # cross-frame queries far from the candidates (warped clouds): bounds stay conservative
    far = (query + torch.tensor([3.0, -2.0, 5.0], device=DEV)).contiguous()
    assert torch.equal(K.knn(far, cand, k), K.knn_bruteforce(far, cand, k))
    if k <= 4:          # one thread per query with a shared tile walk per warp (>= 64 queries per SM) == one warp per query
        from kd_pointcloud_b200 import _lib
        _lib.lib().kdpc_knn_set_few(1)
        try:
            assert torch.equal(K.knn_sorted(qs, cs, b, s, n, k), brute)
        finally:
            _lib.lib().kdpc_knn_set_few(0)


def test_knn_pruned_degenerate_clouds():
    one = torch.ones(2, 1024, 3, device=DEV) * 3.25                      # all points identical
    assert torch.equal(K.knn(one, one, 16), K.knn_bruteforce(one, one, 16))
    line = torch.zeros(1, 4096, 3, device=DEV)
    line[0, :, 0] = torch.arange(4096, device=DEV) * 0.5                  # collinear
    assert torch.equal(K.knn(line, line, 9), K.knn_bruteforce(line, line, 9))
    big = _cloud(1, 2048, 1, "kitti").to(DEV) * 100.0                     # large coordinates: large rounding noise
    assert torch.equal(K.knn(big, big, 32), K.knn_bruteforce(big, big, 32))


def test_knn_batched_and_cached():
    d = make_pairs(3, 1024, seed=9)
    xyz, q = d["pos1"].to(DEV), d["pos2"][:, :200].contiguous().to(DEV)
    KF.clear_caches()
    a = KF.knn_idx(16, xyz, q)
    b = KF.knn_idx(16, xyz, q)
    assert a.data_ptr() == b.data_ptr()                  # second call served from the cache
    assert torch.equal(a.cpu().long(), O.knn_point(16, d["pos1"], q.cpu()))
    q.add_(1.0)                                           # in-place change bumps the version: no stale hit
    c = KF.knn_idx(16, xyz, q)
    assert torch.equal(c.cpu().long(), O.knn_point(16, d["pos1"], q.cpu()))
    assert KF.knn_point(16, xyz, q).dtype == torch.int64


def test_knn_matches_torch_cuda_matmul_topk_chain():
    """The reference op chain itself, run by torch on this GPU (cuBLAS sgemm + topk): index SETS must
    agree except where the cuBLAS dot-product rounding flips an exact K-th boundary."""
    d = make_pairs(2, 4096, seed=17, kind="kitti")
    xyz, q = d["pos1"].to(DEV), d["pos2"].to(DEV)
    torch.backends.cuda.matmul.allow_tf32 = False
    dist = -2 * torch.matmul(q, xyz.permute(0, 2, 1))
    dist += torch.sum(q ** 2, -1).view(2, -1, 1)
    dist += torch.sum(xyz ** 2, -1).view(2, 1, -1)
    ref = torch.sort(torch.topk(dist, 16, dim=-1, largest=False, sorted=False)[1], dim=-1)[0]
    mine = torch.sort(K.knn(q, xyz, 16).long(), dim=-1)[0]
    bad_rows = (ref != mine).any(dim=-1).float().mean().item()
    mine_sq = K.square_distance(q, xyz)
    frac_bits = (mine_sq != dist).float().mean().item()
    print(f"rows differing from torch-CUDA chain: {bad_rows:.2e}; distance entries differing in bits: {frac_bits:.2e}")
    assert bad_rows < 2e-3


def test_square_distance_bit_exact():
    d = make_pairs(2, 700, seed=4, kind="kitti")
    q = d["pos1"][:, :130].contiguous()
    got = K.square_distance(q.to(DEV), d["pos2"].to(DEV)).cpu()
    assert torch.equal(got, O.square_distance_c(q, d["pos2"]))


# ----------------------------------------------------------------------- three_nn / interpolate
@pytest.mark.parametrize("n,m,mode", [(8192, 2048, "ft3d"), (1000, 300, "grid"), (257, 2, "ft3d"), (64, 5000, "dup")])
def test_three_nn_bit_exact(n, m, mode):
    unknown, known = _cloud(2, n, 1, mode), _cloud(2, m, 2, mode)
    d2, idx = K.three_nn(unknown.to(DEV), known.to(DEV))             # the kernel's own outputs: squared
    rd, ri = O.three_nn(unknown, known)
    assert torch.equal(idx.cpu(), ri)
    assert torch.equal(torch.sqrt(d2.cpu()), rd)                       # d2 bit-exact (sqrt taken on the same device)
    dist, idx2 = KF.three_nn(unknown.to(DEV), known.to(DEV))           # public API returns sqrt(d2) (pointnet2_utils.py:98)
    assert torch.equal(idx2, idx) and torch.allclose(dist.cpu(), rd, rtol=1e-6, atol=0, equal_nan=True)


def test_three_interpolate_bit_exact_and_grad():
    g = torch.Generator().manual_seed(0)
    f = torch.randn(2, 19, 300, generator=g)
    idx = torch.randint(0, 300, (2, 1000, 3), generator=g).int()
    w = torch.rand(2, 1000, 3, generator=g)
    fd = f.to(DEV).requires_grad_(True)
    out = KF.three_interpolate(fd, idx.to(DEV), w.to(DEV))
    assert torch.equal(out.detach().cpu(), O.three_interpolate(f, idx, w))
    go = torch.randn(2, 19, 1000, generator=g)
    out.backward(go.to(DEV))
    fr = f.clone().requires_grad_(True)
    ref = (torch.gather(fr.unsqueeze(2).expand(2, 19, 1000, 300), 3, idx.long().unsqueeze(1).expand(2, 19, 1000, 3)) * w.unsqueeze(1)).sum(-1)
    ref.backward(go)
    assert torch.allclose(fd.grad.cpu(), fr.grad, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------- gather / group (cm)
def test_gather_and_group_channel_major_exact_and_grad():
    g = torch.Generator().manual_seed(1)
    f = torch.randn(3, 13, 500, generator=g)
    idx = torch.randint(0, 500, (3, 77), generator=g).int()
    fd = f.to(DEV).requires_grad_(True)
    out = KF.gather_operation(fd, idx.to(DEV))
    assert torch.equal(out.detach().cpu(), O.gather_operation(f, idx))
    go = torch.randn(3, 13, 77, generator=g)
    out.backward(go.to(DEV))
    fr = f.clone().requires_grad_(True)
    O.gather_operation(fr, idx).backward(go)
    assert torch.allclose(fd.grad.cpu(), fr.grad, rtol=1e-5, atol=1e-6)

    for (c, n, s, k) in [(64, 8192, 2048, 16), (3, 1000, 50, 9), (1, 64, 64, 4), (7, 60000, 10, 3)]:
        f = torch.randn(2, c, n, generator=g)
        idx = torch.randint(0, n, (2, s, k), generator=g).int()
        fd = f.to(DEV).requires_grad_(True)
        out = KF.grouping_operation(fd, idx.to(DEV))
        assert torch.equal(out.detach().cpu(), O.grouping_operation(f, idx)), (c, n, s, k)
        if n <= 8192:
            go = torch.randn(2, c, s, k, generator=g)
            out.backward(go.to(DEV))
            fr = f.clone().requires_grad_(True)
            O.grouping_operation(fr, idx).backward(go)
            assert torch.allclose(fd.grad.cpu(), fr.grad, rtol=1e-4, atol=1e-5)


def test_ball_query_exact():
    d = make_pairs(2, 1500, seed=6)
    new_xyz = d["pos1"][:, :333].contiguous()
    for r, ns in ((1.5, 16), (0.05, 8), (100.0, 32)):
        got = KF.ball_query(r, ns, d["pos1"].to(DEV), new_xyz.to(DEV)).cpu()
        assert torch.equal(got, O.ball_query(r, ns, d["pos1"], new_xyz)), (r, ns)


# ----------------------------------------------------------------------- point-major gathers
@pytest.mark.parametrize("c", [3, 64, 131, 20])
def test_gather_rows_and_group_concat_exact(c):
    g = torch.Generator().manual_seed(c)
    d = make_pairs(2, 900, seed=c)
    xyz, q = d["pos1"], d["pos2"][:, :150].contiguous()
    feats = torch.randn(2, 900, c, generator=g)
    idx = torch.randint(0, 900, (2, 150, 16), generator=g).int()
    got = KF.gather_rows(feats.to(DEV), idx.to(DEV)).cpu()
    ref = O.index_points_group(feats, idx)
    assert torch.equal(got, ref)
    fps = torch.randint(0, 900, (2, 40), generator=g).int()
    assert torch.equal(KF.gather_rows(feats.to(DEV), fps.to(DEV)).cpu(), O.index_points_gather(feats, fps))
    gc = KF.group_concat(xyz.to(DEV), q.to(DEV), feats.to(DEV), idx.to(DEV)).cpu()
    rel = O.index_points_group(xyz, idx) - q.view(2, 150, 1, 3)
    assert torch.equal(gc, torch.cat([rel, ref], dim=-1))
    assert torch.equal(KF.group_concat(xyz.to(DEV), q.to(DEV), None, idx.to(DEV)).cpu(), rel)
    # the shared-memory staged kernel (rows * W not a multiple of 4, or >= 2^32 floats) gives the same bits
    from kd_pointcloud_b200 import _lib
    _lib.lib().kdpc_group_concat_set_direct(0)
    try:
        staged = KF.group_concat(xyz.to(DEV), q.to(DEV), feats.to(DEV), idx.to(DEV)).cpu()
    finally:
        _lib.lib().kdpc_group_concat_set_direct(1)
    assert torch.equal(gc, staged)
    idx9 = torch.randint(0, 900, (2, 150, 9), generator=g).int()           # K = 9: pieces straddle rows at every offset
    gc9 = KF.group_concat(xyz.to(DEV), q.to(DEV), feats.to(DEV), idx9.to(DEV)).cpu()
    assert torch.equal(gc9, torch.cat([O.index_points_group(xyz, idx9) - q.view(2, 150, 1, 3), O.index_points_group(feats, idx9)], dim=-1))


def test_group_concat_backward_is_deterministic_and_correct():
    g = torch.Generator().manual_seed(3)
    d = make_pairs(2, 600, seed=8)
    xyz, q = d["pos1"], d["pos2"][:, :128].contiguous()
    feats = torch.randn(2, 600, 24, generator=g)
    idx = O.knn_point(16, xyz, q).int()
    go = torch.randn(2, 128, 16, 27, generator=g)

    def run():
        a, b_, c_ = (t.clone().to(DEV).requires_grad_(True) for t in (xyz, q, feats))
        KF.group_concat(a, b_, c_, idx.to(DEV)).backward(go.to(DEV))
        return a.grad.cpu(), b_.grad.cpu(), c_.grad.cpu()

    g1, g2 = run(), run()
    for x, y in zip(g1, g2):
        assert torch.equal(x, y)                          # bit-identical run to run (no float atomics)
    a, b_, c_ = (t.clone().requires_grad_(True) for t in (xyz, q, feats))
    ref = torch.cat([O.index_points_group(a, idx) - b_.view(2, 128, 1, 3), O.index_points_group(c_, idx)], dim=-1)
    ref.backward(go)
    for x, y in zip(g1, (a.grad, b_.grad, c_.grad)):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-5)


def test_build_csr_inverse_index():
    g = torch.Generator().manual_seed(5)
    idx = torch.randint(0, 300, (3, 70, 9), generator=g).int()
    off, perm = K.build_csr(idx.to(DEV), 300)
    off, perm = off.cpu().numpy(), perm.cpu().numpy()
    flat = idx.reshape(3, -1).numpy()
    for b in range(3):
        assert off[b, 0] == 0 and off[b, -1] == flat.shape[1]
        for i in range(300):
            seg = perm[b, off[b, i]:off[b, i + 1]]
            assert np.array_equal(seg, np.nonzero(flat[b] == i)[0])     # ascending members


def test_knn_on_displaced_clouds_reuses_the_parent_order_with_identical_results():
    """Warped clouds (PointWarping) are sorted by re-using the parent's Morton order (kdpc_spatial_reorder): the kNN
    results must not change - for a smooth displacement, for an incoherent one (loose boxes), and for a permutation-like
    jump - as queries and as candidates."""
    g = torch.Generator().manual_seed(8)
    d = make_pairs(4, 8192, seed=12)
    base, other = d["pos1"].to(DEV), d["pos2"].to(DEV)
    for kind in ("smooth", "noise", "shuffle"):
        if kind == "smooth":
            moved = base + d["flow"].to(DEV)
        elif kind == "noise":
            moved = base + 3.0 * torch.randn(base.shape, generator=g).to(DEV)
        else:
            moved = base[:, torch.randperm(8192, generator=g).to(DEV)].contiguous()
        for k in (3, 32):
            KF.clear_caches()
            ref_a = KF.knn_idx(k, moved, other)                 # moved = candidates
            ref_b = KF.knn_idx(k, other, moved)                 # moved = queries
            KF.clear_caches()
            KF._sorted_cloud(base)                              # the parent is sorted (as in the model)
            KF.hint_displaced_copy(moved, base)
            n0 = K_LAUNCHES()
            got_a = KF.knn_idx(k, moved, other)
            got_b = KF.knn_idx(k, other, moved)
            assert torch.equal(got_a, ref_a) and torch.equal(got_b, ref_b), (kind, k)
    KF.clear_caches()


def K_LAUNCHES():
    from kd_pointcloud_b200 import ops
    return ops.LAUNCHES


def test_build_csr_large_multi_part_and_skewed():
    """Model-size lists (8192 candidates x 32 selections, several CTAs per cloud), a skewed list (every entry selects
    one of 5 candidates: long segments) and candidates nobody selects."""
    g = torch.Generator().manual_seed(6)
    for n, s, k, hi in ((8192, 8192, 32, 8192), (2048, 8192, 3, 2048), (1000, 700, 9, 5), (64, 64, 16, 64), (256, 2048, 12, 256),
                        (512, 4096, 9, 400)):
        idx = torch.randint(0, hi, (2, s, k), generator=g).int()
        off, perm = K.build_csr(idx.to(DEV), n)
        off, perm = off.cpu().numpy().astype(np.int64), perm.cpu().numpy()
        flat = idx.reshape(2, -1).numpy()
        for b in range(2):
            order = np.argsort(flat[b], kind="stable")                  # members of each candidate, ascending
            assert np.array_equal(perm[b], order)
            counts = np.bincount(flat[b], minlength=n)
            assert np.array_equal(off[b], np.concatenate([[0], np.cumsum(counts)]))


# ----------------------------------------------------------------------- fused layer pieces
def test_weightnet_and_aggregation():
    g = torch.Generator().manual_seed(2)
    for wout in (4, 8, 16):
        x = torch.randn(2, 50, 9, 35, generator=g)
        ws = [torch.randn(8, 3, generator=g), torch.randn(8, generator=g), torch.randn(8, 8, generator=g) * 0.5,
              torch.randn(8, generator=g), torch.randn(wout, 8, generator=g) * 0.5, torch.randn(wout, generator=g)]
        got = K.weightnet(x.to(DEV), *[w.to(DEV) for w in ws]).cpu()
        h = torch.relu(x[..., :3] @ ws[0].t() + ws[1])
        h = torch.relu(h @ ws[2].t() + ws[3])
        ref = torch.relu(h @ ws[4].t() + ws[5])
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)         # fp32, different summation order
        grouped = torch.randn(2, 50, 9, 35, generator=g)
        agg = K.pointconv_agg(grouped.to(DEV), got.to(DEV)).cpu()
        ref_agg = torch.matmul(grouped.permute(0, 1, 3, 2), got).reshape(2, 50, -1)   # pointconv_util.py:249
        assert torch.allclose(agg, ref_agg, rtol=1e-5, atol=1e-5)


def test_costvol_pre_and_max_over_k():
    g = torch.Generator().manual_seed(4)
    d = make_pairs(2, 400, seed=10)
    x1, x2 = d["pos1"][:, :100].contiguous(), d["pos2"]
    D = 32
    p1, p2 = torch.randn(2, 100, D, generator=g), torch.randn(2, 400, D, generator=g)
    idx = O.knn_point(16, x2, x1).int()
    pw, pb = torch.randn(D, 3, generator=g), torch.randn(D, generator=g)
    got = K.costvol_pre(*[t.to(DEV) for t in (x1, x2, p1, p2, idx, pw, pb)], 0.1).cpu()
    rel = O.index_points_group(x2, idx) - x1.view(2, 100, 1, 3)
    ref = torch.nn.functional.leaky_relu(O.index_points_group(p2, idx) + p1.view(2, 100, 1, D) + rel @ pw.t() + pb, 0.1)
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5)
    mx, arg = K.max_over_k(got.to(DEV))
    rm, ra = ref.max(dim=2)
    assert torch.allclose(mx.cpu(), rm, rtol=1e-5, atol=1e-5)
    assert torch.equal(torch.gather(got, 2, arg.cpu().long().unsqueeze(2)).squeeze(2), mx.cpu())


@pytest.mark.parametrize("c", [3, 64, 30])
def test_interp3_matches_oracle_and_grad(c):
    g = torch.Generator().manual_seed(c)
    d = make_pairs(2, 800, seed=12)
    dense = d["pos1"]
    sparse = dense[:, ::4].contiguous()
    feat = torch.randn(2, 200, c, generator=g)
    idx = KF.knn_idx(3, sparse.to(DEV), dense.to(DEV))
    fd = feat.to(DEV).requires_grad_(True)
    out = KF.interp3(dense.to(DEV), sparse.to(DEV), idx, fd)
    ref = O.upsample_flow(dense.permute(0, 2, 1), sparse.permute(0, 2, 1), feat.permute(0, 2, 1)).permute(0, 2, 1)
    assert torch.allclose(out.detach().cpu(), ref, rtol=1e-5, atol=1e-6)    # tolerance: fp32 weights, 1 ulp sqrt/div
    comp = KF.interp3_composite(dense.to(DEV), sparse.to(DEV), idx, feat.to(DEV))
    assert torch.allclose(comp.cpu(), ref, rtol=1e-5, atol=1e-6)
    go = torch.randn(2, 800, c, generator=g)
    out.backward(go.to(DEV))
    fr = feat.clone().requires_grad_(True)
    O.upsample_flow(dense.permute(0, 2, 1), sparse.permute(0, 2, 1), fr.permute(0, 2, 1)).permute(0, 2, 1).backward(go)
    assert torch.allclose(fd.grad.cpu(), fr.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("B,S,Kn,C", [(2, 300, 9, 131), (1, 77, 16, 67), (2, 64, 9, 515), (1, 5, 3, 4)])
def test_pointconv_agg_grad_matches_fp64(B, S, Kn, C):
    """Backward of the PointConv aggregation (one kernel for both gradients) against the two fp64 bmm of autograd."""
    from kd_pointcloud_b200 import functional as KF
    torch.manual_seed(S + C)
    dev = "cuda:0"
    grouped = torch.randn(B, S, Kn, C, device=dev, requires_grad=True)
    wn = torch.rand(B, S, Kn, 16, device=dev, requires_grad=True)
    go = torch.randn(B, S, C * 16, device=dev)
    out = KF.pointconv_agg(grouped, wn)
    out.backward(go)
    gd, wd = grouped.detach().double().requires_grad_(True), wn.detach().double().requires_grad_(True)
    ref = torch.einsum("bskc,bskw->bscw", gd, wd).reshape(B, S, C * 16)
    ref.backward(go.double())
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    assert rel(out, ref) < 1e-5 and rel(grouped.grad, gd.grad) < 1e-5 and rel(wn.grad, wd.grad) < 1e-5
    # only one of the two gradients requested; deterministic
    g2 = grouped.detach().clone().requires_grad_(True)
    KF.pointconv_agg(g2, wn.detach()).backward(go)
    assert torch.equal(g2.grad, grouped.grad)
    w2 = wn.detach().clone().requires_grad_(True)
    KF.pointconv_agg(grouped.detach(), w2).backward(go)
    assert torch.equal(w2.grad, wn.grad)


@pytest.mark.gpu
def test_square_distance_is_differentiable_like_the_reference():
    from kd_pointcloud_b200 import functional as KF
    torch.manual_seed(3)
    src = (torch.rand(2, 50, 3, device="cuda:0") * 10).requires_grad_(True)
    dst = (torch.rand(2, 70, 3, device="cuda:0") * 10).requires_grad_(True)
    d = KF.square_distance(src, dst)
    assert torch.equal(d.detach(), KF.square_distance(src.detach(), dst.detach()))      # the kernel's values
    w = torch.randn_like(d)
    (d * w).sum().backward()
    s2, d2 = src.detach().double().requires_grad_(True), dst.detach().double().requires_grad_(True)
    ref = ((s2.unsqueeze(2) - d2.unsqueeze(1)) ** 2).sum(-1)
    (ref * w.double()).sum().backward()
    assert ((src.grad.double() - s2.grad).abs().max() / s2.grad.abs().max()).item() < 1e-4
    assert ((dst.grad.double() - d2.grad).abs().max() / d2.grad.abs().max()).item() < 1e-4


def test_concat_rows_equals_torch_cat_for_contiguous_and_strided_parts():
    """kdpc_concat_rows (the estimator's [features | upsampled features | cost volume] tensor): bit-identical to
    torch.cat for contiguous parts, column blocks and batch halves of wider tensors, 1..4 parts, odd row counts."""
    from kd_pointcloud_b200 import ops
    g = torch.Generator().manual_seed(11)
    wide = torch.randn(6, 333, 96, generator=g).to(DEV)
    a = wide[:3, :, :32]                      # batch half + column block (row stride 96)
    b = torch.randn(3, 333, 64, generator=g).to(DEV)
    c = wide[3:, :, 64:]                      # second half, last column block
    d = torch.randn(3, 333, 4, generator=g).to(DEV)
    for parts in ((a,), (a, b), (a, b, c), (a, b, c, d), (d, c, b, a)):
        got = ops.concat_rows(parts)
        assert got.is_contiguous() and torch.equal(got, torch.cat(parts, dim=2))
    dest = torch.full((3, 333, 200), -1.0, device=DEV)
    ops.concat_rows((a, b), out=dest[:, :, 100:196])
    assert torch.equal(dest[:, :, 100:196], torch.cat((a, b), dim=2))
    assert bool((dest[:, :, :100] == -1).all()) and bool((dest[:, :, 196:] == -1).all())
    assert ops.concat_rows((torch.empty(0, 5, 8, device=DEV), torch.empty(0, 5, 4, device=DEV))).shape == (0, 5, 12)
    with pytest.raises(ValueError):
        ops.concat_rows((a, torch.randn(3, 333, 6, device=DEV)))           # width % 4 != 0
    with pytest.raises(ValueError):
        ops.concat_rows((a, b.permute(0, 2, 1)[:, :333, :64]))             # not row-strided
