"""GPU edge cases: empty inputs (device-side row count 0 under a non-zero capacity) leave outputs untouched and terminate."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def H():
    from lidiff_b200 import _lib
    return _lib.get_handle(DEV)


@pytest.mark.parametrize("cin,cout", [(32, 32), (128, 128), (256, 256), (3, 32)])
def test_convolution_with_zero_live_rows_is_a_noop(cin, cout):
    from lidiff_b200 import _lib
    from lidiff_b200._lib import ConvDesc, ConvIO
    h = H()
    cap, kvol = 1000, 27
    W = torch.randn(kvol, cin, cout, device=DEV) * 0.05
    Wp = h.pack_weights(W) if cin % 16 == 0 else None
    x = torch.randn(cap, cin, device=DEV)
    xh = torch.zeros(cap, 2 * cin, dtype=torch.float16, device=DEV)
    out = torch.full((cap, cout), 7.0, device=DEV)
    out_h = torch.full((cap, 2 * cout), 3.0, dtype=torch.float16, device=DEV)
    nbr = torch.randint(0, cap, (kvol, cap), dtype=torch.int32, device=DEV)
    mask = torch.full((cap,), (1 << 27) - 1, dtype=torch.int32, device=DEV)
    d_m = torch.zeros(1, dtype=torch.int32, device=DEV)
    for algo in ((_lib.ALGO_FFMA, _lib.ALGO_TC, _lib.ALGO_TC_TILE) if Wp is not None else (_lib.ALGO_FFMA,)):
        d = ConvDesc()
        d.c1, d.c2, d.cout, d.kvol = cin, 0, cout, kvol
        d.weight, d.weight_packed = W.data_ptr(), (Wp.data_ptr() if Wp is not None else None)
        d.relu = 1
        d.nbr, d.nbr_stride, d.d_mout, d.mout_cap, d.npass = nbr.data_ptr(), cap, d_m.data_ptr(), cap, 1
        d.row_mask = mask.data_ptr()
        d.io[0] = ConvIO(x.data_ptr(), None, None, out.data_ptr(), None, None, None, None,
                         xh.data_ptr() if cin % 16 == 0 else None, None, out_h.data_ptr() if cout % 4 == 0 else None, None)
        h.spconv(d, algo)
        torch.cuda.synchronize()
        assert bool((out == 7.0).all()) and bool((out_h == 3.0).all()), f"algo {algo} wrote rows beyond the live count"


def test_maps_order_and_matching_with_zero_rows():
    from lidiff_b200.engine import Geometry
    h = H()
    coords = torch.cat([torch.zeros(500, 1), torch.randint(-50, 50, (500, 3)).float()], 1).to(DEV).contiguous()
    g = Geometry(h, 500, with_up=False)
    g.build(coords, 500)
    zero = torch.zeros(1, dtype=torch.int32, device=DEV)
    nbr = torch.full((27, 500), 5, dtype=torch.int32, device=DEV)
    mask = torch.full((500,), 9, dtype=torch.int32, device=DEV)
    h.kernel_map(g.grid[0], g.C[0], zero, 500, 3, 1, nbr, 500, None, mask)
    assert bool((nbr == -1).all()) and bool((mask == 0).all())
    perm = torch.full((500,), -7, dtype=torch.int32, device=DEV)
    scratch = torch.zeros((h.row_order_scratch_bytes(500) + 3) // 4, dtype=torch.int32, device=DEV)
    for kvol in (27, 8):
        h.row_order(mask, zero, 500, kvol, perm, scratch)
        assert bool((perm == -7).all())
    idx = torch.full((500,), -3, dtype=torch.int32, device=DEV)
    tree = h.nn_tree(g.C[4], g.d_n[4], 500)
    h.nn_match_tree(g.C[0], zero, 500, tree, 500, idx)
    assert bool((idx == -3).all())
    tree0 = h.nn_tree(g.C[4], zero, 500)                      # no keys at all: every query answers 0 like the exhaustive kernel
    a, b = torch.empty(500, dtype=torch.int32, device=DEV), torch.empty(500, dtype=torch.int32, device=DEV)
    h.nn_match_tree(g.C[0], g.d_n[0], 500, tree0, 500, a)
    h.nn_match(g.C[0], g.d_n[0], 500, g.C[4], zero, 500, 0, b)
    n0 = int(g.d_n[0])
    assert torch.equal(a[:n0], b[:n0])
