import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
from doubleattentionspeakerverification_b200 import ops
from oracle import path_oracle as po
from test_gpu_frontend import CASES, bf16_round, dev

for (B, T, F, Cin, Cout, pool, ref, with_len) in CASES + [(1, 20, 80, 64, 128, False, False, False), (2, 16, 80, 64, 64, True, False, False), (1, 12, 20, 64, 128, False, False, False), (1, 12, 20, 64, 128, True, False, False)]:
    rs = np.random.RandomState(B * 100 + T + Cin)
    x = bf16_round(np.maximum(rs.standard_normal((B, T, F, Cin)), 0).astype(np.float32))
    w = bf16_round((rs.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32))
    bias = (rs.standard_normal((Cout,)) * 0.1).astype(np.float32)
    lengths = None
    if with_len:
        lengths = rs.randint(1, T + 1, size=(B,)).astype(np.int32); lengths[0] = T
        x = po._zero_rows(x, lengths)
    ref_y = po._zero_rows(po.relu(po.conv3x3_same(x, w, bias)), lengths)
    if pool:
        ref_y = po.maxpool2x2_ceil(ref_y)
        if ref:
            Bq, T2, F2, C = ref_y.shape
            ref_y = ref_y.transpose(0, 1, 3, 2).reshape(Bq, T2, C * F2)
    y = ops.conv3x3_igemm_bf16(dev(x, torch.bfloat16), ops.pack_conv_weight_bf16(dev(w)), dev(bias), Cout,
                               lengths=None if lengths is None else dev(lengths), pool=pool, ref_layout=ref,
                               out_dtype=torch.float32).float().cpu().numpy()
    err = np.abs(y - ref_y)
    bad = np.argwhere(err > 0.02 * np.abs(ref_y).max())
    print((B, T, F, Cin, Cout, pool, ref, with_len), 'max_rel %.3e' % (err.max() / np.abs(ref_y).max()), 'nbad', len(bad), 'of', err.size,
          'first bad', bad[:3].tolist(), 'last bad', bad[-2:].tolist(), flush=True)
