"""GPU parity of the VGG front-end kernels: fp32 CUDA-core path (1e-4) and the bf16 tcgen05
implicit-GEMM path (bf16 tolerances), against the numpy oracle and the live-reference fixtures."""
import numpy as np
import pytest
import torch

from conftest import golden, max_rel, min_cosine
from doubleattentionspeakerverification_b200 import CNNs, ops, synth
from oracle import path_oracle as po

pytestmark = pytest.mark.gpu


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


def bf16_round(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


@pytest.mark.parametrize('idx', range(4))
def test_frontend_fp32_golden(idx):
    g = golden('frontend_%d.npz' % idx)
    front = str(g['front'])
    K, B, T, seed = [int(v) for v in g['spec']]
    cfg = synth.example_config(front_end=front, kernel_size=K, embedding_size=32, heads_number=8, num_spkrs=4)
    sd = synth.make_state_dict(cfg, seed)
    net = (CNNs.VGG3L if front == 'VGG3L' else CNNs.VGG4L)(K, precision='fp32')
    synth.load_state_dict(net, sd, 'front_end.')
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(dev(synth.make_logmel(B, T, seed)))
    assert y.shape == g['out'].shape
    assert max_rel(y.cpu().numpy(), g['out']) < 1e-4


def test_frontend_bf16_golden_k512():
    g = golden('frontend_3.npz')
    K, B, T, seed = [int(v) for v in g['spec']]
    cfg = synth.example_config(kernel_size=K, embedding_size=32, heads_number=8, num_spkrs=4)
    net = synth.load_state_dict(CNNs.VGG4L(K, precision='bf16'), synth.make_state_dict(cfg, seed), 'front_end.').cuda().eval()
    assert net.resolved_precision() == 'bf16'
    with torch.no_grad():
        y = net(dev(synth.make_logmel(B, T, seed)))
    assert y.dtype == torch.float32 and y.shape == g['out'].shape
    assert min_cosine(y.cpu().numpy().reshape(B, -1), g['out'].reshape(B, -1)) > 0.9999
    assert max_rel(y.cpu().numpy(), g['out']) < 3e-2


CASES = [  # B, T, F, Cin, Cout, pool, ref, with_lengths
    (1, 16, 80, 64, 64, True, False, False),
    (2, 20, 80, 64, 128, False, False, False),
    (2, 21, 40, 128, 128, True, False, True),
    (3, 13, 20, 128, 256, False, False, True),
    (3, 12, 20, 256, 256, True, False, False),
    (5, 7, 10, 64, 192, False, False, True),
    (5, 7, 10, 128, 264, True, True, True),
    (4, 50, 10, 64, 128, True, True, False),
    (1, 3, 80, 64, 8, True, False, False),
    # the deep layers of the exampleModel config (conv32 / conv41 / conv42: K = 9*Cin = 4608 / 9216)
    (2, 12, 20, 512, 512, True, False, True),
    (3, 25, 10, 512, 1024, False, False, True),
    (3, 26, 10, 1024, 1024, True, True, True),
    (2, 9, 10, 1024, 136, True, True, False),
]


PAIR_CASES = [  # Cout multiple of 256: eligible for CTA pairs (cta_group::2)
    (3, 13, 20, 128, 256, False, False, True),
    (3, 12, 20, 256, 256, True, False, False),
    (4, 50, 10, 64, 256, True, True, False),
    (5, 7, 10, 128, 512, True, True, True),
    (2, 30, 40, 64, 256, True, False, True),
    (7, 9, 10, 64, 256, False, False, True),
    (2, 12, 20, 512, 512, True, False, True),
    (3, 25, 10, 512, 1024, False, False, True),
    (3, 26, 10, 1024, 1024, True, True, True),
    (4, 50, 10, 1024, 256, True, True, False),
]


@pytest.mark.parametrize('B,T,F,Cin,Cout,pool,ref,with_len', PAIR_CASES)
def test_igemm_cta_pairs_vs_oracle(B, T, F, Cin, Cout, pool, ref, with_len):
    test_igemm_vs_oracle(B, T, F, Cin, Cout, pool, ref, with_len, pair=True)


@pytest.mark.parametrize('B,T,F,Cin,Cout,pool,ref,with_len', CASES)
def test_igemm_vs_oracle(B, T, F, Cin, Cout, pool, ref, with_len, pair=False):
    rs = np.random.RandomState(B * 100 + T + Cin)
    x = bf16_round(np.maximum(rs.standard_normal((B, T, F, Cin)), 0).astype(np.float32))
    w = bf16_round((rs.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32))
    bias = (rs.standard_normal((Cout,)) * 0.1).astype(np.float32)
    lengths = None
    if with_len:
        lengths = rs.randint(1, T + 1, size=(B,)).astype(np.int32)
        lengths[0] = T
        x = po._zero_rows(x, lengths)          # the layer's input already obeys the masking rule
    ref_y = po._zero_rows(po.relu(po.conv3x3_same(x, w, bias)), lengths)
    if pool:
        ref_y = po.maxpool2x2_ceil(ref_y)
        if ref:
            Bq, T2, F2, C = ref_y.shape
            ref_y = ref_y.transpose(0, 1, 3, 2).reshape(Bq, T2, C * F2)
    y = ops.conv3x3_igemm_bf16(dev(x, torch.bfloat16), ops.pack_conv_weight_bf16(dev(w)), dev(bias), Cout,
                               lengths=None if lengths is None else dev(lengths), pool=pool, ref_layout=ref,
                               out_dtype=torch.float32, pair=pair)
    assert tuple(y.shape) == ref_y.shape
    # inputs are bf16-exact and accumulation is fp32: only the summation order and, for the NHWC
    # bf16 outputs, the final rounding to bf16 (2^-9 relative) differ from the oracle
    tol = 1e-4 if y.dtype == torch.float32 else 6e-3
    assert max_rel(y.float().cpu().numpy(), ref_y) < tol


def _round_to(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).float().numpy()


@pytest.mark.parametrize('xdt,wdt', [(torch.float16, torch.float16)])
@pytest.mark.parametrize('B,T,F,Cin,Cout,pool,ref,pair', [(2, 21, 40, 128, 128, True, False, False), (3, 13, 20, 128, 256, False, False, True),
                                                         (3, 26, 10, 512, 512, True, True, True), (2, 9, 10, 64, 136, True, True, False)])
def test_igemm_operand_formats(xdt, wdt, B, T, F, Cin, Cout, pool, ref, pair):
    """fp16 operands (precision='fp16').  Inputs exactly representable in fp16, fp32 accumulation: only the summation order
    and the output rounding (2^-12) differ from the oracle.  A mixed bf16/fp16 pair is an error, not a fault."""
    with pytest.raises(Exception):
        ops.conv3x3_igemm_bf16(torch.zeros(1, 4, 10, 64, device='cuda', dtype=torch.bfloat16),
                               ops.pack_conv_weight_bf16(torch.zeros(64, 64, 3, 3, device='cuda'), torch.float16),
                               torch.zeros(64, device='cuda'), 64)
    rs = np.random.RandomState(7 + B + Cin)
    x = _round_to(np.maximum(rs.standard_normal((B, T, F, Cin)), 0).astype(np.float32), xdt)
    w = _round_to((rs.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32), wdt)
    bias = (rs.standard_normal((Cout,)) * 0.1).astype(np.float32)
    lengths = rs.randint(1, T + 1, size=(B,)).astype(np.int32)
    lengths[0] = T
    x = po._zero_rows(x, lengths)
    ref_y = po._zero_rows(po.relu(po.conv3x3_same(x, w, bias)), lengths)
    if pool:
        ref_y = po.maxpool2x2_ceil(ref_y)
        if ref:
            Bq, T2, F2, C = ref_y.shape
            ref_y = ref_y.transpose(0, 1, 3, 2).reshape(Bq, T2, C * F2)
    y = ops.conv3x3_igemm_bf16(dev(x, xdt), ops.pack_conv_weight_bf16(dev(w), wdt), dev(bias), Cout, lengths=dev(lengths),
                               pool=pool, ref_layout=ref, out_dtype=torch.float32, pair=pair)
    assert tuple(y.shape) == ref_y.shape and y.dtype == (torch.float32 if ref else xdt)
    tol = 1e-4 if ref else (6e-3 if xdt == torch.bfloat16 else 8e-4)
    assert max_rel(y.float().cpu().numpy(), ref_y) < tol


@pytest.mark.parametrize('B,T,F,Cin,Cout,pool,ref,pair', [(2, 21, 40, 128, 128, True, False, False), (3, 13, 20, 128, 256, False, False, True),
                                                         (3, 26, 10, 512, 512, True, True, True), (2, 9, 10, 64, 136, True, True, False),
                                                         (2, 12, 20, 256, 256, True, False, True)])
def test_igemm_fp32x3_vs_oracle(B, T, F, Cin, Cout, pool, ref, pair):
    """fp32x3 mode: arbitrary fp32 activations and weights as bf16 hi/lo pairs, three MMAs per product, against the fp32
    oracle: 3e-5 (the dropped lo*lo term and the pairs' own rounding are ~2^-16 per product).  NHWC outputs come back split."""
    rs = np.random.RandomState(17 + B + Cin)
    x = np.maximum(rs.standard_normal((B, T, F, Cin)), 0).astype(np.float32)
    w = (rs.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32)
    bias = (rs.standard_normal((Cout,)) * 0.1).astype(np.float32)
    lengths = rs.randint(1, T + 1, size=(B,)).astype(np.int32)
    lengths[0] = T
    x = po._zero_rows(x, lengths)
    ref_y = po._zero_rows(po.relu(po.conv3x3_same(x, w, bias)), lengths)
    if pool:
        ref_y = po.maxpool2x2_ceil(ref_y)
        if ref:
            Bq, T2, F2, C = ref_y.shape
            ref_y = ref_y.transpose(0, 1, 3, 2).reshape(Bq, T2, C * F2)
    xt = dev(x)
    hi = xt.to(torch.bfloat16)
    xs = torch.cat([hi, (xt - hi.float()).to(torch.bfloat16)], dim=-1).contiguous()
    y = ops.conv3x3_igemm_bf16(xs, ops.pack_conv_weight_x3(dev(w)), dev(bias), Cout, lengths=dev(lengths), pool=pool, ref_layout=ref,
                               out_dtype=torch.float32, pair=pair, x3=True)
    if not ref:
        assert y.dtype == torch.bfloat16 and y.shape[-1] == 2 * Cout
        y = y[..., :Cout].float() + y[..., Cout:].float()
    assert tuple(y.shape) == ref_y.shape
    assert max_rel(y.cpu().numpy(), ref_y) < 3e-5


def test_conv11_split_output():
    rs = np.random.RandomState(6)
    x = (2 * rs.standard_normal((2, 9, 80))).astype(np.float32)
    w = rs.standard_normal((16, 1, 3, 3)).astype(np.float32)
    b = rs.standard_normal((16,)).astype(np.float32)
    L = np.array([9, 4], np.int32)
    want = ops.conv11_direct(dev(x), dev(w), dev(b), dev(L))
    y = ops.conv11_direct(dev(x), dev(w), dev(b), dev(L), split=True)
    assert y.dtype == torch.bfloat16 and tuple(y.shape) == (2, 9, 80, 32)
    got = y[..., :16].float() + y[..., 16:].float()
    assert float((got - want).abs().max()) <= 2 ** -15 * float(want.abs().max())
    assert torch.equal(y[..., :16], want.to(torch.bfloat16))


@pytest.mark.parametrize('B,T,F,C1,Cout,pool,act,with_len', [(2, 37, 80, 128, 128, True, torch.bfloat16, True),
                                                             (3, 16, 80, 64, 64, True, torch.bfloat16, False),
                                                             (1, 5, 80, 128, 128, True, torch.float16, False),
                                                             (4, 50, 40, 128, 256, False, torch.bfloat16, True),
                                                             (2, 401, 80, 128, 128, True, torch.bfloat16, True),
                                                             (9, 23, 20, 64, 136, True, torch.float16, True)])
def test_conv12_fused_equals_separate_kernels(B, T, F, C1, Cout, pool, act, with_len):
    """conv11 computed inside conv12's kernel (per-tile scratch patches in L2) == conv11_direct followed by the implicit
    GEMM, bit for bit, including masked rows, image borders and ragged lengths."""
    rs = np.random.RandomState(B * 7 + T)
    x = dev((2 * rs.standard_normal((B, T, F))).astype(np.float32))
    w11 = dev((rs.standard_normal((C1, 1, 3, 3)) * 0.5).astype(np.float32))
    b11 = dev((rs.standard_normal((C1,)) * 0.1).astype(np.float32))
    w12 = dev((rs.standard_normal((Cout, C1, 3, 3)) * np.sqrt(2.0 / (9 * C1))).astype(np.float32))
    b12 = dev((rs.standard_normal((Cout,)) * 0.1).astype(np.float32))
    L = None
    if with_len:
        Ln = rs.randint(1, T + 1, size=(B,)).astype(np.int32)
        Ln[0] = T
        L = dev(Ln)
    wp = ops.pack_conv_weight_bf16(w12, act)
    h = ops.conv11_direct(x, w11, b11, L, out_dtype=act)
    want = ops.conv3x3_igemm_bf16(h, wp, b12, Cout, L, pool=pool)
    for _ in range(2):                                             # twice: the scratch buffers are reused
        got = ops.conv12_fused(x, w11, b11, wp, b12, Cout, L, pool=pool, act_dtype=act)
        assert got.dtype == act and got.shape == want.shape
        assert torch.equal(got, want)


@pytest.mark.parametrize('B,T,F,Cin,Cout,pool,ref,dt', [(1, 50, 10, 1024, 1024, True, True, torch.bfloat16),
                                                        (1, 50, 10, 512, 1024, False, False, torch.bfloat16),
                                                        (1, 100, 20, 256, 512, False, False, torch.float16),
                                                        (2, 101, 20, 512, 512, True, False, torch.bfloat16),
                                                        (1, 37, 40, 256, 264, True, False, torch.bfloat16),
                                                        (4, 7, 10, 1024, 1024, True, True, torch.bfloat16),      # several utterances per patch
                                                        (3, 5, 20, 512, 512, False, False, torch.bfloat16),
                                                        (5, 9, 4, 1024, 256, True, False, torch.float16)])
def test_igemm_split_k_small_batches(B, T, F, Cin, Cout, pool, ref, dt):
    """Launches with too few tiles for the SMs run split along K (partials through a workspace + a finishing kernel):
    same results as the oracle, and the shapes here really take that path."""
    from doubleattentionspeakerverification_b200 import _lib
    pair = pool and Cin >= 256 and Cout % 256 == 0                  # as CNNs.py asks for the pooled layers
    flags = ops.CONV_RELU | (ops.CONV_POOL if pool else 0) | (ops.CONV_REF_LAYOUT if ref else 0) | (ops.CONV_PAIR if pair else 0) | \
        ((ops.CONV_W_F16 | ops.CONV_X_F16) if dt == torch.float16 else 0)
    ydt = 0 if ref else (2 if dt == torch.float16 else 1)
    nws = _lib.lib().dasv_conv3x3_igemm_workspace_bytes(ydt, flags, 0, B, T, F, Cin, Cout)
    assert nws % (B * T * F * Cout * 4) == 0
    if Cin >= 512 and B == 1:
        assert nws > 0, 'the deep layers at batch 1 are expected to run split along K'
    rs = np.random.RandomState(3 + T + Cin)
    x = _round_to(np.maximum(rs.standard_normal((B, T, F, Cin)), 0).astype(np.float32), dt)
    w = _round_to((rs.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32), dt)
    bias = (rs.standard_normal((Cout,)) * 0.1).astype(np.float32)
    ref_y = po.relu(po.conv3x3_same(x, w, bias))
    if pool:
        ref_y = po.maxpool2x2_ceil(ref_y)
        if ref:
            Bq, T2, F2, C = ref_y.shape
            ref_y = ref_y.transpose(0, 1, 3, 2).reshape(Bq, T2, C * F2)
    y = ops.conv3x3_igemm_bf16(dev(x, dt), ops.pack_conv_weight_bf16(dev(w), dt), dev(bias), Cout, pool=pool, ref_layout=ref,
                               out_dtype=torch.float32, pair=pair)
    assert tuple(y.shape) == ref_y.shape
    tol = 1e-4 if ref else (6e-3 if dt == torch.bfloat16 else 8e-4)
    assert max_rel(y.float().cpu().numpy(), ref_y) < tol
    y2 = ops.conv3x3_igemm_bf16(dev(x, dt), ops.pack_conv_weight_bf16(dev(w), dt), dev(bias), Cout, pool=pool, ref_layout=ref,
                                out_dtype=torch.float32, pair=pair)
    assert torch.equal(y, y2)                                       # fixed summation order: deterministic


def test_fp16_store_saturates():
    """fp16 activations saturate at the largest finite value instead of overflowing to infinity."""
    x = torch.full((1, 4, 80), 3.0e4, device='cuda')
    w = torch.ones((8, 1, 3, 3), device='cuda')
    y = ops.conv11_direct(x, w, torch.zeros(8, device='cuda'), None, out_dtype=torch.float16)
    assert bool(torch.isfinite(y).all()) and float(y.max()) == 65504.0


def test_conv11_and_pool_kernels():
    rs = np.random.RandomState(5)
    x = rs.standard_normal((2, 9, 80)).astype(np.float32)
    w = rs.standard_normal((16, 1, 3, 3)).astype(np.float32)
    b = rs.standard_normal((16,)).astype(np.float32)
    L = np.array([9, 4], np.int32)
    ref = po._zero_rows(po.relu(po.conv3x3_same(po._zero_rows(x[..., None], L), w, b)), L)
    y = ops.conv11_direct(dev(x), dev(w), dev(b), dev(L))
    assert max_rel(y.cpu().numpy(), ref) < 1e-5
    yb = ops.conv11_direct(dev(x), dev(w), dev(b), dev(L), out_dtype=torch.bfloat16)
    assert max_rel(yb.float().cpu().numpy(), ref) < 6e-3
    p = ops.maxpool2x2(y)
    pp = po.maxpool2x2_ceil(y.cpu().numpy())          # pooling itself is exact
    assert max_rel(p.cpu().numpy(), pp) == 0.0
    pr = ops.maxpool2x2(y, ref_layout=True)
    assert max_rel(pr.cpu().numpy(), pp.transpose(0, 1, 3, 2).reshape(2, 5, -1)) == 0.0


@pytest.mark.parametrize('T,F,Cin,Cout,pool', [(64, 20, 128, 128, True), (64, 20, 128, 256, False), (50, 10, 512, 512, True), (37, 8, 64, 128, False)])
def test_lazy_masking_leaves_valid_rows_and_the_halo_row_intact(T, F, Cin, Cout, pool):
    """DASV_CONV_LAZY_MASK: tiles wholly beyond an utterance's length are skipped.  Rows below the length are bit-identical to the
    eager-masking result and the one row the next 3x3 layer reads beyond them is zero -- also when the rows further down
    of the INPUT hold garbage (NaN here), which is what a lazily masked previous layer leaves behind."""
    lens = np.array([T, T - 1, (3 * T) // 4, T // 2, T // 2 - 1, 13, 12, 2, 1, 0], np.int32)
    B = len(lens)
    rs = np.random.RandomState(T + Cin)
    x = np.maximum(rs.standard_normal((B, T, F, Cin)), 0).astype(np.float32)
    w = (rs.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32)
    bias = (rs.standard_normal((Cout,)) * 0.1).astype(np.float32)
    x_clean, x_dirty = x.copy(), x.copy()
    for b, L in enumerate(lens):
        x_clean[b, L:] = 0
        x_dirty[b, L:] = 0
        x_dirty[b, L + 1:] = np.nan                               # row L is the zero halo; anything below is garbage
    wp, bd, Ld = ops.pack_conv_weight_bf16(dev(w)), dev(bias), dev(lens)
    want = ops.conv3x3_igemm_bf16(dev(x_clean, torch.bfloat16), wp, bd, Cout, Ld, pool=pool)
    got = ops.conv3x3_igemm_bf16(dev(x_dirty, torch.bfloat16), wp, bd, Cout, Ld, pool=pool, lazy_mask=True)
    for b, L in enumerate(lens):
        Lo = (L + 1) // 2 if pool else L
        assert torch.equal(got[b, :Lo], want[b, :Lo]), b
        if Lo < got.shape[1]:
            assert not got[b, Lo].float().abs().max().item() > 0, b   # the row the next layer's last valid row looks at
        assert torch.isfinite(got[b, :min(Lo + 1, got.shape[1])].float()).all()


def test_lazy_masking_first_layer():
    rs = np.random.RandomState(5)
    T, F = 70, 80
    lens = np.array([70, 69, 33, 32, 31, 8, 1, 0], np.int32)
    x = rs.standard_normal((len(lens), T, F)).astype(np.float32)
    w = rs.standard_normal((128, 1, 3, 3)).astype(np.float32) * 0.3
    bias = rs.standard_normal((128,)).astype(np.float32) * 0.1
    want = ops.conv11_direct(dev(x), dev(w), dev(bias), dev(lens), out_dtype=torch.bfloat16)
    got = ops.conv11_direct(dev(x), dev(w), dev(bias), dev(lens), out_dtype=torch.bfloat16, lazy_mask=True)
    for b, L in enumerate(lens):
        assert torch.equal(got[b, :L], want[b, :L])
        if L < T:
            assert not got[b, L].float().abs().max().item() > 0
