"""Host-side logic of the data-parallel extractor (no GPU): sharding plan, batching/padding, and the
world_size-2 all-gather order restoration over gloo with a stub embedder."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from doubleattentionspeakerverification_b200 import extract


def stub_embed(x, lengths):
    """Deterministic per-utterance 'embedding' that depends only on the valid frames (so padding or batch
    composition errors change it): [sum, sum of squares, length, first-frame mean]."""
    out = []
    for b in range(x.shape[0]):
        v = x[b, :int(lengths[b])].double()
        out.append(torch.stack([v.sum(), (v * v).sum(), torch.tensor(float(lengths[b]), dtype=torch.float64), v[0].mean()]))
    return torch.stack(out).float()


def make_feats(n, seed=0):
    rs = np.random.RandomState(seed)
    return [rs.standard_normal((int(rs.randint(20, 200)), 8)).astype(np.float32) for _ in range(n)]


def test_shard_plan_is_balanced_partition():
    lengths = np.random.RandomState(1).randint(200, 2000, size=101)
    for world in (1, 2, 4, 8):
        plan = extract.shard_plan(lengths, world)
        allidx = np.sort(np.concatenate(plan))
        assert np.array_equal(allidx, np.arange(101))
        sizes = [len(p) for p in plan]
        assert max(sizes) - min(sizes) <= 1
        frames = [lengths[p].sum() for p in plan]
        assert max(frames) / min(frames) < 1.1


def test_global_plan_partitions_balances_and_respects_budget():
    """BASELINE configs[3] lengths (2-20 s): every utterance in exactly one batch, batches within the frame budget (the
    short-utterance batches may pass it by up to 50 % in padded frames), rank loads within a few % of each other, and far fewer small batches than
    per-rank bucketing."""
    for world in (1, 2, 4, 8):
        rs = np.random.RandomState(0)
        L = (100 * rs.uniform(2, 20, size=256 * world)).astype(np.int64)
        gp = extract.global_plan(L, world)
        allidx = np.sort(np.concatenate([b for bs in gp for b in bs]))
        assert np.array_equal(allidx, np.arange(len(L)))
        for bs in gp:
            for b in bs:
                assert len(b) * L[b].max() <= 1.5 * 256 * 400
        loads = [sum(extract._batch_cost(L, b) for b in bs) for bs in gp]
        assert max(loads) <= 1.05 * np.mean(loads)
        small = sum(1 for bs in gp for b in bs if len(b) * L[b].max() < 16 * 400)
        old_small = sum(1 for p in extract.shard_plan(L, world) for b in extract.bucket_plan(L[p], 256 * 400, 0.8) if len(b) * L[p][b].max() < 16 * 400)
        assert small <= 1 and old_small >= 2
    # degenerate inputs
    assert extract.global_plan(np.array([5]), 4) == [[np.array([0])], [], [], []] or sum(len(bs) for bs in extract.global_plan(np.array([5]), 4)) == 1
    assert sum(len(bs) for bs in extract.global_plan(np.zeros((0,), np.int64), 2)) == 0


def test_batch_plan_respects_budget_and_covers_all():
    lengths = np.random.RandomState(2).randint(200, 2000, size=57)
    batches = extract.batch_plan(lengths, max_frames=8000, max_batch=6)
    assert np.array_equal(np.sort(np.concatenate(batches)), np.arange(57))
    for b in batches:
        assert len(b) <= 6 and (len(b) == 1 or len(b) * lengths[b].max() <= 8000)
    x, L = extract.pad_batch(make_feats(5), [4, 0, 2])
    assert x.shape[0] == 3 and x.shape[1] == L.max() and np.all(x[1, L[1]:] == 0)


def test_packed_path_and_bucket_plan():
    lengths = np.random.RandomState(6).randint(200, 2000, size=64)
    batches = extract.bucket_plan(lengths, max_frames=20000, min_ratio=0.8)
    assert np.array_equal(np.sort(np.concatenate(batches)), np.arange(64))
    for b in batches:
        assert lengths[b].min() >= 0.8 * lengths[b].max() and (len(b) == 1 or len(b) * lengths[b].max() <= 20000)
    feats = make_feats(19, seed=7)
    packed = extract.PackedUtterances(feats, pin=False)
    emb = extract.extract_sharded(stub_embed, packed, 'cpu', max_frames=700)
    want = torch.cat([stub_embed(torch.from_numpy(f)[None], torch.tensor([f.shape[0]])) for f in feats])
    assert torch.allclose(emb, want, rtol=1e-5, atol=1e-5)


def test_sparse_packed_holds_only_owned_utterances():
    feats = make_feats(9, seed=11)
    L = [f.shape[0] for f in feats]
    owned = {i: feats[i] for i in (1, 4, 8)}
    sp = extract.PackedUtterances.sparse(L, owned, pin=False)
    assert len(sp) == 9 and sp.data.shape[0] == sum(L[i] for i in owned)
    emb = extract.extract_local_packed(stub_embed, sp, [8, 1, 4], 'cpu', max_frames=10 ** 6)
    want = torch.cat([stub_embed(torch.from_numpy(feats[i])[None], torch.tensor([L[i]])) for i in (8, 1, 4)])
    assert torch.allclose(emb, want, rtol=1e-5, atol=1e-5)


def test_single_process_matches_per_utterance():
    feats = make_feats(23, seed=3)
    emb = extract.extract_sharded(stub_embed, feats, 'cpu', max_frames=600)
    want = torch.cat([stub_embed(torch.from_numpy(f)[None], torch.tensor([f.shape[0]])) for f in feats])
    assert torch.allclose(emb, want, rtol=1e-5, atol=1e-5)


def _worker(rank, world, port, n):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        feats = make_feats(n, seed=4)
        emb = extract.extract_sharded(stub_embed, feats, 'cpu', max_frames=500, embedding_size=4)
        want = torch.cat([stub_embed(torch.from_numpy(f)[None], torch.tensor([f.shape[0]])) for f in feats])
        assert emb.shape == want.shape
        assert torch.allclose(emb, want, rtol=1e-5, atol=1e-5), 'rank %d: gathered embeddings out of order' % rank
    finally:
        dist.destroy_process_group()


def _validate_worker(rank, world, port):
    """validate(): every rank scores its slice of the trial lists; the threshold counts are summed over the ranks, so the
    EER equals the single-process sweep over all scores."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from doubleattentionspeakerverification_b200 import utils
        feats = make_feats(13, seed=8)
        rs = np.random.RandomState(1)
        cl = np.stack([rs.randint(0, 13, 41), rs.randint(0, 13, 41)], 1)
        im = np.stack([rs.randint(0, 13, 30), rs.randint(0, 13, 30)], 1)

        def score(emb, trials):
            t = torch.as_tensor(trials)
            return torch.nn.functional.cosine_similarity(emb[t[:, 0]], emb[t[:, 1]], dim=-1)

        def count(scores):
            return (scores.double()[None, :] >= torch.from_numpy(utils.EER_THRESHOLDS)[:, None]).sum(1)

        eer, CL, IM = extract.validate(stub_embed, feats, cl, im, 'cpu', embedding_size=4, max_frames=500, score_fn=score, count_fn=count)
        want = torch.cat([stub_embed(torch.from_numpy(f)[None], torch.tensor([f.shape[0]])) for f in feats])
        ge = [count(score(want, t)).numpy() for t in (cl, im)]
        assert eer == utils.eer_from_counts(ge[0], len(cl), ge[1], len(im))
        lo, hi = extract.rank_slice(len(cl), rank, world)
        assert CL.numel() == hi - lo and torch.allclose(CL, score(want, cl[lo:hi]), atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_validate_slices_trials_and_sums_counts():
    mp.spawn(_validate_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_rank_slices_cover_rows():
    for n in (0, 1, 7, 1024):
        for world in (1, 2, 3, 8):
            sl = [extract.rank_slice(n, r, world) for r in range(world)]
            assert sl[0][0] == 0 and sl[-1][1] == n and all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
            assert max(h - l for l, h in sl) - min(h - l for l, h in sl) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world2_gloo_allgather_restores_order():
    for n in (11, 1):      # odd count (uneven shards) and fewer utterances than ranks (an empty shard)
        mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)


def _train_worker(rank, world, port):
    """Two ranks, uneven shards of one batch: all-reduced gradients == the single-process gradient of the batch-mean loss."""
    from doubleattentionspeakerverification_b200 import train_utils
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        if rank == 1:
            for p in net.parameters():
                p.data.add_(1.0)                                 # ranks start out different ...
        train_utils.broadcast_parameters(net)                      # ... until rank 0's values are broadcast
        x = torch.randn(7, 6, generator=torch.Generator().manual_seed(1))
        y = torch.randn(7, 3, generator=torch.Generator().manual_seed(2))
        sl = train_utils.shard_batch(7)
        assert (sl.start, sl.stop) == ((0, 4) if rank == 0 else (4, 7))
        loss = ((net(x[sl]) - y[sl]) ** 2).sum(1).mean()
        loss.backward()
        train_utils.allreduce_gradients(list(net.parameters()), bucket_bytes=64, local_weight=(sl.stop - sl.start) / 7.0)
        got = [p.grad.clone() for p in net.parameters()]
        net.zero_grad()
        ((net(x) - y) ** 2).sum(1).mean().backward()
        for g, p in zip(got, net.parameters()):
            assert torch.allclose(g, p.grad, rtol=1e-5, atol=1e-6)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_gradient_allreduce_equals_global_batch():
    mp.spawn(_train_worker, args=(2, _free_port()), nprocs=2, join=True)


def stub_features(wave, n_samples, sfr):
    """Stand-in for the GPU log-mel kernels in the host-logic tests: 'frames' of 100 samples, 4 moments per frame."""
    B = wave.shape[0]
    frames = np.asarray(n_samples) // 100
    T = int(frames.max())
    x = wave[:, :T * 100].reshape(B, T, 100)
    feat = torch.stack([x.mean(2), x.abs().mean(2), x.min(2).values, x.max(2).values], dim=2)
    return feat, torch.from_numpy(frames.astype(np.int32))


def _audio_worker(rank, world, port):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(3)
        waves = [rs.standard_normal(int(rs.randint(600, 5000))).astype(np.float32) for _ in range(9)]
        emb = extract.extract_sharded_audio(stub_embed, waves, 16000, 'cpu', embedding_size=4, feature_fn=stub_features, max_samples=9000)
        for i, w in enumerate(waves):
            f, fr = stub_features(torch.from_numpy(w)[None], [len(w)], 16000)
            assert torch.allclose(emb[i], stub_embed(f, fr)[0], rtol=1e-5, atol=1e-5), 'rank %d utterance %d' % (rank, i)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_waveform_extraction_restores_order():
    mp.spawn(_audio_worker, args=(2, _free_port()), nprocs=2, join=True)
