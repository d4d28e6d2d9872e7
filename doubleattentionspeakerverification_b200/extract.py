"""Batched, variable-length, data-parallel embedding extraction and trial scoring.

Replaces the reference's validation loop (scripts/train.py:117-133: two batch-1 forwards and one
``.item()`` sync per trial line) with: every unique utterance embedded once, in padded + length-masked
batches (SURVEY.md §5.7 masking rule => identical to per-utterance batch-1 results), utterances sharded
over the ranks of one ``torchrun`` job (one process per GPU), ONE all-gather of the ``[N/R, E]`` embedding
shards, then batched cosine scoring.  Utterances are independent, so there is no other collective.

The host logic here (sharding plan, batching, order restoration) is device-agnostic and covered by
world_size-2 ``gloo`` tests on CPU with a stub embedder; the product embedder is
``SpeakerClassifier.getEmbedding`` on CUDA.
"""
import numpy as np
import torch

from . import utils


def shard_plan(lengths, world):
    """Length-sorted snake deal: sort utterances by decreasing length and deal them to ranks 0..R-1, R-1..0,
    ..., so every rank gets the same number (±1) and nearly the same total frames.  Returns ``[index arrays]``
    per rank (each sorted by decreasing length).  Deterministic: every rank computes the same plan."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind='stable')
    pos = np.arange(len(order))
    rnd, col = pos // world, pos % world
    owner = np.where(rnd % 2 == 0, col, world - 1 - col)
    return [order[owner == r] for r in range(world)]


def batch_plan(lengths, max_frames, max_batch=None, multiple=1):
    """Greedy batches over a length-sorted index list: a batch's padded size (its longest utterance rounded
    up to ``multiple`` x its count) stays <= max_frames.  Returns a list of index arrays."""
    idx = np.argsort(-np.asarray(lengths), kind='stable')
    batches, cur = [], []
    for i in idx:
        Tpad = -(-int(lengths[cur[0]] if cur else lengths[i]) // multiple) * multiple
        if cur and ((len(cur) + 1) * Tpad > max_frames or (max_batch and len(cur) >= max_batch)):
            batches.append(np.array(cur))
            cur = []
        cur.append(int(i))
    if cur:
        batches.append(np.array(cur))
    return batches


class PackedUtterances:
    """Every utterance's frames in ONE (pinned) host buffer ``[sum T_i, F]`` plus offsets -- the form a feature loader
    should hand over: no per-batch host padding, the device builds padded batches itself with one gather."""

    def __init__(self, feats, pin=True):
        self.lengths = np.array([f.shape[0] for f in feats], np.int64)
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)])
        F = feats[0].shape[1]
        self.data = torch.empty((int(self.offsets[-1]), F), dtype=torch.float32)
        if pin and torch.cuda.is_available():
            self.data = self.data.pin_memory()
        view = self.data.numpy()
        for f, o in zip(feats, self.offsets[:-1]):
            view[o:o + f.shape[0]] = f

    def __len__(self):
        return len(self.lengths)


def bucket_plan(lengths, max_frames, min_ratio=0.8, max_batch=None):
    """Batches over the length-sorted list whose padded size stays <= max_frames AND whose shortest utterance is at
    least ``min_ratio`` of the longest, which bounds the padding waste of a batch by 1/min_ratio - 1."""
    idx = np.argsort(-np.asarray(lengths), kind='stable')
    batches, cur = [], []
    for i in idx:
        if cur:
            Tmax = int(lengths[cur[0]])
            if (len(cur) + 1) * Tmax > max_frames or lengths[i] < min_ratio * Tmax or (max_batch and len(cur) >= max_batch):
                batches.append(np.array(cur))
                cur = []
        cur.append(int(i))
    if cur:
        batches.append(np.array(cur))
    return batches


def extract_local_packed(embed_fn, packed, indices, device, max_frames=256 * 400, min_ratio=0.8, max_batch=None):
    """Embed ``packed`` utterances ``indices``: their frames go to the device once (one async copy per utterance from
    the pinned buffer, no host padding), then every batch is ONE device gather into ``[B, Tmax, F]`` with frame indices
    clamped to the utterance (the kernels ignore frames >= length, so the padding content does not matter).
    Returns ``[len(indices), E]`` in the order of ``indices``."""
    indices = np.asarray(indices)
    if len(indices) == 0:
        return None
    dev = torch.device(device)
    L = packed.lengths[indices]
    starts = np.concatenate([[0], np.cumsum(L)])
    frames = torch.empty((int(starts[-1]), packed.data.shape[1]), device=dev, dtype=torch.float32)
    for j, i in enumerate(indices):
        o = int(packed.offsets[i])
        frames[int(starts[j]):int(starts[j + 1])].copy_(packed.data[o:o + int(L[j])], non_blocking=True)
    starts_d = torch.from_numpy(starts[:-1]).to(dev)
    L_d = torch.from_numpy(L).to(dev)
    out = None
    for b in bucket_plan(L, max_frames, min_ratio, max_batch):
        bt = torch.from_numpy(b).to(dev)
        Tmax = int(L[b].max())
        t = torch.arange(Tmax, device=dev)
        rows = starts_d[bt, None] + torch.minimum(t[None, :], L_d[bt, None] - 1)          # [B, Tmax] source frame of every slot
        x = frames[rows]                                                                   # one gather: padded batch
        emb = embed_fn(x, L_d[bt].to(torch.int32))
        if out is None:
            out = torch.empty((len(indices), emb.shape[1]), device=emb.device, dtype=emb.dtype)
        out[bt] = emb
    return out


def extract_sharded_audio(embed_fn, waves, sfr, device, group=None, embedding_size=None, feature_fn=None, **kw):
    """``extract_sharded`` for waveforms: ranks take length-balanced shards of ``waves`` (1-D float arrays), compute
    features and embeddings on their GPU (``extract_local_audio``), then one all-gather restores ``[N, E]`` in the
    original order on every rank."""
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    plan = shard_plan(np.array([len(w) for w in waves]), world)
    local = extract_local_audio(embed_fn, waves, plan[rank], sfr, device, feature_fn=feature_fn, **kw)
    return _gather_shards(local, plan, len(waves), device, group, embedding_size)


def extract_local_audio(embed_fn, waves, indices, sfr, device, max_samples=256 * 64352, max_batch=None, min_ratio=0.8,
                        feature_fn=None, **feature_kw):
    """Waveform -> embedding without the host in between (getEmbeddingExample.py:22-36 = extractFeatures + getEmbedding):
    ``waves[i]`` (1-D float arrays in [-1, 1)) go to the device in zero-padded batches of similar length, the log-mel
    features + CMN are computed there (``featureExtractor.logmel_batch``) and handed to ``embed_fn(x [B,T,80], frames
    [B])`` with the frame counts as lengths.  Returns ``[len(indices), E]`` in the order of ``indices``."""
    indices = np.asarray(indices)
    if len(indices) == 0:
        return None
    if feature_fn is None:                                   # (wave [B,N], n_samples, sfr) -> (features [B,T,F], frames [B])
        from . import featureExtractor as fe
        feature_fn = fe.logmel_batch
    dev = torch.device(device)
    n = np.array([len(waves[i]) for i in indices])
    if (n < 512).any():
        raise ValueError('a waveform is shorter than one analysis frame (512 samples)')
    out = None
    for b in bucket_plan(n, max_samples, min_ratio, max_batch):
        nb = n[b]
        host = torch.zeros((len(b), int(nb.max())), dtype=torch.float32, pin_memory=dev.type == 'cuda')
        for j, i in enumerate(b):
            host[j, :nb[j]] = torch.from_numpy(np.asarray(waves[indices[i]], dtype=np.float32))
        feat, frames = feature_fn(host.to(dev, non_blocking=True), nb, sfr, **feature_kw)
        emb = embed_fn(feat, frames)
        if out is None:
            out = torch.empty((len(indices), emb.shape[1]), device=emb.device, dtype=emb.dtype)
        out[torch.from_numpy(b).to(emb.device)] = emb
    return out


def pad_batch(feats, idx, multiple=1):
    """Stack ``feats[i]`` (``[T_i, F]`` arrays) into a zero-padded ``[len(idx), Tmax, F]`` float32 array + lengths."""
    L = np.array([feats[i].shape[0] for i in idx], np.int32)
    Tmax = -(-int(L.max()) // multiple) * multiple
    x = np.zeros((len(idx), Tmax, feats[idx[0]].shape[1]), np.float32)
    for j, i in enumerate(idx):
        x[j, :L[j]] = feats[i]
    return x, L


def extract_local(embed_fn, feats, indices, device, max_frames=256 * 400, max_batch=None):
    """Embed ``feats[i] for i in indices`` in padded, length-masked batches.  ``embed_fn(x [B,T,F], lengths [B])
    -> [B,E]`` (e.g. ``net.getEmbedding``).  Returns ``[len(indices), E]`` in the order of ``indices``."""
    indices = np.asarray(indices)
    if len(indices) == 0:
        return None
    lengths = np.array([feats[i].shape[0] for i in indices])
    out = None
    for b in batch_plan(lengths, max_frames, max_batch):
        x, L = pad_batch(feats, indices[b])
        xt = torch.from_numpy(x)
        if torch.device(device).type == 'cuda':
            xt = xt.pin_memory().to(device, non_blocking=True)
        emb = embed_fn(xt, torch.from_numpy(L).to(device))
        if out is None:
            out = torch.empty((len(indices), emb.shape[1]), device=emb.device, dtype=emb.dtype)
        out[torch.from_numpy(b).to(emb.device)] = emb
    return out


def extract_sharded(embed_fn, feats, device, group=None, max_frames=256 * 400, max_batch=None, embedding_size=None, min_ratio=0.8):
    """Every rank embeds its shard, then one all-gather; returns ``[N, E]`` embeddings of ALL utterances, in
    the original order, on every rank.  Without an initialised process group it is the single-GPU path.
    ``feats``: list of ``[T_i, F]`` arrays (host-padded batches) or a ``PackedUtterances`` (device-built batches)."""
    import torch.distributed as dist
    N = len(feats)
    is_packed = isinstance(feats, PackedUtterances)
    lengths = feats.lengths if is_packed else np.array([f.shape[0] for f in feats])
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size(group) if distributed else 1
    rank = dist.get_rank(group) if distributed else 0
    plan = shard_plan(lengths, world)
    if is_packed:
        local = extract_local_packed(embed_fn, feats, plan[rank], device, max_frames, min_ratio=min_ratio, max_batch=max_batch)
    else:
        local = extract_local(embed_fn, feats, plan[rank], device, max_frames, max_batch)
    return _gather_shards(local, plan, N, device, group, embedding_size)


def _gather_shards(local, plan, N, device, group, embedding_size):
    """All-gather the ranks' ``[len(plan[rank]), E]`` shards and restore the original utterance order."""
    import torch.distributed as dist
    world = len(plan)
    if world == 1:
        out = torch.empty_like(local)
        out[torch.from_numpy(plan[0]).to(local.device)] = local
        return out
    per = max(len(p) for p in plan)                       # shards differ by at most one utterance: pad to equal
    E = local.shape[1] if local is not None else embedding_size
    if E is None:
        raise ValueError('a rank with an empty shard needs embedding_size')
    send = torch.zeros((per, E), device=device, dtype=torch.float32)
    if local is not None:
        send[:local.shape[0]] = local
    recv = torch.empty((world * per, E), device=device, dtype=torch.float32)
    dist.all_gather_into_tensor(recv, send, group=group)   # the path's only collective (NCCL over NVLink on GPUs)
    out = torch.empty((N, E), device=device, dtype=torch.float32)
    for r in range(world):
        n = len(plan[r])
        if n:
            out[torch.from_numpy(plan[r]).to(device)] = recv[r * per:r * per + n]
    return out


class HostPipeline:
    """Double-buffered host -> device -> host streaming of fixed-shape batches: the H2D copy of batch i+1 runs on a
    copy stream while batch i is being embedded on the compute stream, and the embeddings return to pinned host
    memory asynchronously.  Usage:  pipe = HostPipeline(embed_fn, (B, T, F), E, device);  pipe.submit(x_pinned, out_pinned)
    per batch, then pipe.finish().  ``embed_fn(x_dev) -> [B, E]``; an optional ``post(emb)`` hook runs on the compute
    stream between the extraction and the D2H copy (e.g. the NCCL all-gather)."""

    def __init__(self, embed_fn, shape, embedding_size, device, post=None):
        self.embed_fn, self.post, self.device = embed_fn, post, torch.device(device)
        self.stage = [torch.empty(shape, device=self.device, dtype=torch.float32) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.i = 0

    def submit(self, x_host, out_host):
        k = self.i & 1
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.copy_stream):
            if self.i >= 2:
                self.copy_stream.wait_event(self.consumed[k])      # the staging buffer's previous batch has been read
            self.stage[k].copy_(x_host, non_blocking=True)
            self.copied[k].record(self.copy_stream)
        compute.wait_event(self.copied[k])
        emb = self.embed_fn(self.stage[k])
        self.consumed[k].record(compute)
        if self.post is not None:
            emb = self.post(emb)
        out_host.copy_(emb, non_blocking=True)
        self.i += 1
        return emb

    def finish(self):
        torch.cuda.current_stream(self.device).synchronize()


def score_trial_list(emb, trials, device=None):
    """``trials``: ``[M,2]`` integer array of (utterance A, utterance B) indices -> ``[M]`` cosine scores."""
    t = torch.as_tensor(np.asarray(trials), device=emb.device if device is None else device)
    return utils.score_pairs(emb, t[:, 0], t[:, 1])


def score_cross(emb, enrol_idx, test_idx):
    """Cross-product scoring: ``[len(enrol_idx), len(test_idx)]`` cosine matrix."""
    e = emb[torch.as_tensor(np.asarray(enrol_idx), device=emb.device)].contiguous()
    t = emb[torch.as_tensor(np.asarray(test_idx), device=emb.device)].contiguous()
    return utils.score_matrix(e, t)


def validate(embed_fn, utterances, client_trials, impostor_trials, device, **kw):
    """The reference's validation pass (scripts/train.py:158-184: __extract_scores on the client and impostor
    trial lists, then __calculate_EER) without its per-trial forwards: every utterance is embedded once (sharded over
    the ranks when a process group is up), both trial lists are scored by one kernel each, and the 200-threshold
    FAR/FRR sweep runs on the device.  ``*_trials``: ``[M,2]`` utterance-index pairs.  Returns (EER, CL, IM)."""
    emb = extract_sharded(embed_fn, utterances, device, **kw)
    CL = score_trial_list(emb, client_trials)
    IM = score_trial_list(emb, impostor_trials)
    return utils.calculate_EER(CL, IM), CL, IM
