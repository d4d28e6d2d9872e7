#!/usr/bin/env python
"""Headline benchmark: speaker-embedding extraction throughput (embeddings/sec, 4 s utterances) of the
exampleModel config (VGG4L K=1024 + DoubleMHA H=32 + FC/BN, E=400) on synthetic log-mel, plus the
DoubleMHA pooling microbench (BASELINE.json configs[1]) as achieved HBM GB/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (getEmbedding) over one batch of 256 utterances x 400 frames x 80
bins per GPU (weak scaling: utterances are independent, sharded over ranks, no data-path collective).
``value`` is timed with inputs resident in HBM; ``e2e`` through the public module API with pinned HOST
input, H2D + D2H inside the timed region (and, for N > 1, the NCCL all-gather of the embeddings).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# keep stdout to the single JSON line: NCCL's version / debug banner goes to stderr
# stdout carries ONLY the JSON line(s): everything else any library prints (e.g. NCCL's version banner, which goes to
# the C-level stdout) is sent to stderr by pointing fd 1 at fd 2 and keeping the real stdout aside for the result
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (line + '\n').encode())


import numpy as np   # noqa: E402
import torch         # noqa: E402

BATCH, FRAMES, BINS = 256, 400, 80
METRIC, UNIT = 'embeddings/sec (4 s utts)', 'embeddings/s'


def workload_config(world, batch=BATCH):
    """The `config` object of the JSON line: the same keys and values from both arms (--impl b200 / reference)."""
    return {'workload': 'full embedding extraction VGG4L(K=1024)+DoubleMHA(H=32)+FC(E=400), 4 s (400x80) log-mel '
                        'utterances, random-init (BASELINE configs[2])', 'batch_per_gpu': batch,
            'parallelism': 'dp%d' % world,
            'l2': 'two rotating input batches; per-step intermediates (>3 GB) exceed the 126 MB L2'}


def profiled_traffic():
    """DRAM bytes per conv launch from the committed ncu capture of this round (profiles/r2_conv_ncu.json, written by
    scripts/ncu_summary.py from `ncu --set full`): dram__bytes_read.sum + dram__bytes_write.sum averaged over the
    seven launches of a step.  None when the file is missing, so a stale literal can never be reported."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_conv_ncu.json')) as f:
            d = json.load(f)
        return float(d['dram_bytes_per_launch']), d.get('source', 'profiles/r2_conv_ncu.json')
    except Exception:
        return None, 'no ncu capture committed for this round'


def peaks():
    p = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            m = json.load(f)
        p.update(hbm_gbs=float(m['hbm_gbs']), bf16_tflops=float(m['bf16_tflops']),
                 bf16_tflops_sustained=float(m.get('bf16_tflops_sustained', m['bf16_tflops'])), source='measured')
    except Exception:
        pass
    return p


def conv_flops(batch, frames, kernel_size=1024):
    """Algorithmic FLOPs of the VGG4L conv stack (2*MACs), per layer name."""
    from doubleattentionspeakerverification_b200 import synth
    out, T, F = {}, frames, BINS
    for i, ((cin, cout), name) in enumerate(zip(synth.vgg_channels('VGG4L', kernel_size), synth.conv_names('VGG4L'))):
        out[name] = 2.0 * batch * T * F * cout * 9 * cin
        if i % 2 == 1:
            T, F = (T + 1) // 2, (F + 1) // 2
    return out


def conv_io_bytes(batch, frames, kernel_size=1024):
    """Compulsory HBM bytes of every tensor-core conv launch: 16-bit NHWC input + output (fp32 for the last layer) + weights."""
    from doubleattentionspeakerverification_b200 import synth
    out, T, F = {}, frames, BINS
    chans = synth.vgg_channels('VGG4L', kernel_size)
    names = synth.conv_names('VGG4L')
    for i, ((cin, cout), name) in enumerate(zip(chans, names)):
        pooled = i % 2 == 1
        To, Fo = ((T + 1) // 2, (F + 1) // 2) if pooled else (T, F)
        if i > 0:
            out[name] = batch * T * F * cin * 2.0 + batch * To * Fo * cout * (4.0 if i == len(names) - 1 else 2.0) + 9.0 * cin * cout * 2
        if pooled:
            T, F = To, Fo
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed regions run."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_rate(target_s, threads=None):
    """The reference's CPU path (oracle/torch_port.py: the same torch calls as scripts/model.py:52-59 on the
    reference's layouts), exampleModel config, 4 s utterances, all host threads.  Bounded sample."""
    from doubleattentionspeakerverification_b200 import synth
    from oracle import torch_port as tp
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = synth.example_config()
    sd = tp.as_torch(synth.make_state_dict(cfg, 1234))
    nb = 4
    x = torch.from_numpy(synth.make_logmel(nb, FRAMES, 7))
    tp.get_embedding(x[:1], sd, cfg)                      # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        tp.get_embedding(x, sd, cfg)
        n += nb
        el = time.perf_counter() - t0
        if el >= target_s or n >= BATCH:
            break
    return n / el, n, el, threads


def run_reference(args, rank):
    if rank != 0:
        return
    per_step = []
    total_budget = 60.0
    step_s = max(1.0, total_budget / max(1, args.steps + args.warmup))
    for i in range(args.warmup + args.steps):
        rate, n, el, threads = cpu_reference_rate(step_s)
        if i >= args.warmup:
            per_step.append((n, el))
    n_tot = sum(n for n, _ in per_step)
    t_tot = sum(t for _, t in per_step)
    value = n_tot / t_tot
    sample = '%d utterances of the 256 x 400-frame batch per step (bounded CPU sample), fp32, %d torch threads' % (
        per_step[0][0], threads)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / len(per_step), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.gpus),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp16', 'fp32x3', 'fp32'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-dmha', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary measurements (training step, feature extraction, other precisions)')
    ap.add_argument('--pairs', action='store_true', help='A/B: pooled layers with >= 256 input channels on CTA pairs (cta_group::2), the default until late in round 2')
    ap.add_argument('--no-configs', action='store_true', help='skip BASELINE configs[3] (2-20 s ragged) and configs[4] (1 M trials)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl b200 needs a CUDA device (there is no CPU fallback)')
    from doubleattentionspeakerverification_b200 import _lib, extract, model, ops, synth
    _lib.lib()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    pk = peaks()

    cfg = synth.example_config()
    cfg.precision = args.precision
    net = synth.load_state_dict(model.SpeakerClassifier(cfg, dev), synth.make_state_dict(cfg, 1234)).to(dev).eval()
    net.front_end.use_pairs = bool(args.pairs)
    Bn = args.batch
    xs = [torch.from_numpy(synth.make_logmel(Bn, FRAMES, seed=100 + rank * 7 + i)).to(dev) for i in range(2)]
    x_host = torch.from_numpy(synth.make_logmel(Bn, FRAMES, seed=300 + rank)).pin_memory()
    emb_host = torch.empty((Bn, cfg.embedding_size), dtype=torch.float32).pin_memory()
    gathered = torch.empty((world * Bn, cfg.embedding_size), device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident(i):
        with torch.no_grad():
            return net.getEmbedding(xs[i & 1])

    def post(emb):
        if world > 1:
            dist.all_gather_into_tensor(gathered, emb)           # the path's only collective (cosine scoring needs all embeddings)
        return emb

    def embed(xd):
        with torch.no_grad():
            return net.getEmbedding(xd)

    pipe = extract.HostPipeline(embed, (Bn, FRAMES, BINS), cfg.embedding_size, dev, post=post)

    def step_e2e(i):
        # public API: pinned host batch in, pinned host embeddings out; the H2D of step i+1 overlaps step i's kernels
        pipe.submit(x_host, emb_host)

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if drain is not None:
            drain()                                   # the timed region ends when the last D2H copy / collective has finished
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        return max_over_ranks(ms)

    for i in range(args.warmup):
        step_resident(i)
    for i in range(2):
        step_e2e(i)
    with ClockSampler(local) as clk:                 # clocks / throttle reasons over both timed regions
        _lib.LAUNCHES.clear()
        ms = timed(step_resident, args.steps)
        launches = sum(_lib.LAUNCHES.values())
        ms_e2e = timed(step_e2e, args.steps, drain=lambda: torch.cuda.current_stream().wait_stream(pipe.out_stream))
    value = world * Bn * args.steps / (ms * 1e-3)
    e2e_value = world * Bn * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (conv3x3_igemm: tensor-bound), timed per launch inside real steps
    roof = None
    if args.precision in ('bf16', 'fp16'):
        fl = conv_flops(Bn, FRAMES)
        names = [n for n in synth.conv_names('VGG4L')][1:]
        rec = []
        orig = ops.conv3x3_igemm_bf16

        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y = orig(*a, **k)
            e1.record()
            rec.append((e0, e1))
            return y

        nsteps = max(3, min(args.steps, 10))
        graphs, net.use_graphs = net.use_graphs, False      # per-launch events need the eager launches (same kernels as the graph replays)
        for i in range(2):                                  # eager warm-up: the first eager step after the graph replays allocates
            step_resident(i)                                # its activations anew, and that host stall would land inside conv12's events
        torch.cuda.synchronize()
        ops.conv3x3_igemm_bf16 = wrapped
        for i in range(nsteps):
            step_resident(i)
        torch.cuda.synchronize()
        net.use_graphs = graphs
        ops.conv3x3_igemm_bf16 = orig
        per_layer = {n: 0.0 for n in names}
        for j, (e0, e1) in enumerate(rec):
            per_layer[names[j % len(names)]] += e0.elapsed_time(e1) / nsteps
        conv_ms = sum(per_layer.values())
        conv_fl = sum(fl[n] for n in names)
        achieved = conv_fl / (conv_ms * 1e-3) / 1e12
        roof = {'bound': 'tensor', 'kernel': 'conv3x3_igemm_kernel (7 launches per step, conv12..conv42)',
                'achieved': achieved, 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / pk['bf16_tflops_sustained'], 'peak_source': pk['source'] + ' sustained cuBLAS bf16',
                'traffic': profiled_traffic()[0], 'traffic_source': profiled_traffic()[1],
                'algorithmic_bytes_per_launch': sum(conv_io_bytes(Bn, FRAMES).values()) / len(names),
                'avg_launch_ms': conv_ms / len(names), 'share_of_step': conv_ms / (ms / args.steps),
                'per_layer_tflops': {n: fl[n] / (per_layer[n] * 1e-3) / 1e12 for n in names}}

    # ---- BASELINE configs[3] (2-20 s utterances, packed + masked, sharded) and configs[4] (1 M trials) at this N
    cfg3 = cfg4 = None
    if not args.no_configs:
        def embed_len(xb, L):
            with torch.no_grad():
                return net.getEmbedding(xb, lengths=L)

        def timed_call(fn, reps=2):
            fn()                                                  # warm-up
            best = None
            for _ in range(reps):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn()
                e1.record()
                torch.cuda.synchronize()
                t = max_over_ranks(e0.elapsed_time(e1))
                best = t if best is None else min(best, t)
            return best * 1e-3, out

        per_gpu = 256
        N3 = per_gpu * world
        rs3 = np.random.RandomState(0)
        frames3 = (100 * rs3.uniform(2.0, 20.0, size=N3)).astype(np.int64)
        base = synth.make_logmel(1, 2000, seed=1)[0]
        gp = extract.global_plan(frames3, world)
        mine = set(int(i) for b in gp[rank] for i in b)
        # every rank holds the frames of ITS utterances only (the others are never touched: the plan decides the owner)
        packed3 = extract.PackedUtterances.sparse(frames3, {i: np.roll(base, i * 7, axis=0)[:int(frames3[i])] for i in sorted(mine)})
        dt3, _ = timed_call(lambda: extract.extract_sharded(embed_len, packed3, dev, embedding_size=cfg.embedding_size))
        pad = sum(len(b) * int(frames3[b].max()) for bs in gp for b in bs)
        loads = [sum(extract._batch_cost(frames3, b) for b in bs) for bs in gp]
        useful = 12.99e9 * frames3.sum() / 100.0                  # conv FLOPs of the VALID frames (12.99 GFLOP per second of audio)
        cfg3 = {'workload': 'BASELINE configs[3]: %d utterances/GPU, durations U[2,20] s, one global batch plan (extract.global_plan), '
                            'padded + length-masked batches, one all-gather' % per_gpu,
                'utterances': int(N3), 'seconds': dt3, 'embeddings_per_s': N3 / dt3,
                'useful_conv_tflops_per_gpu': useful / dt3 / 1e12 / world, 'frac_of_sustained_peak': useful / dt3 / 1e12 / world / pk['bf16_tflops_sustained'],
                'padding_waste': float(pad / frames3.sum() - 1.0), 'batches': int(sum(len(bs) for bs in gp)),
                'plan_imbalance': float(max(loads) / (sum(loads) / world)),
                'timing': 'CUDA events, max over ranks, best of 2: pinned host frames -> H2D -> device gather -> extraction -> all-gather'}
        del packed3

        M4 = 2048
        lo4, hi4 = extract.rank_slice(M4, rank, world)
        x4 = torch.from_numpy(np.stack([np.roll(base, int(i) * 3, axis=0)[:FRAMES] for i in range(lo4, hi4)])).pin_memory()
        per4 = -(-M4 // world)
        gath4 = torch.empty((world * per4, cfg.embedding_size), device=dev)

        def trials():
            embs = []
            for j in range(0, x4.shape[0], Bn):                  # this rank's share of the 2 048 utterances, 256 at a time
                with torch.no_grad():
                    embs.append(net.getEmbedding(x4[j:j + Bn].to(dev, non_blocking=True)))
            local = torch.cat(embs)
            if world > 1:
                send = torch.zeros((per4, cfg.embedding_size), device=dev)
                send[:local.shape[0]] = local
                dist.all_gather_into_tensor(gath4, send)          # NCCL over NVLink: the path's only collective
                emb_all = torch.cat([gath4[r * per4:r * per4 + (extract.rank_slice(M4, r, world)[1] - extract.rank_slice(M4, r, world)[0])] for r in range(world)])
            else:
                emb_all = local
            return extract.score_cross_sharded(emb_all, np.arange(1024), np.arange(1024, 2048), gather=True)

        dt4, scores4 = timed_call(trials)
        cfg4 = {'workload': 'BASELINE configs[4]: 1024 x 1024 = 1 048 576 trials: 2 048 4 s utterances extracted across the ranks from pinned host '
                            'memory, all-gather of the embeddings, enrol rows of the cosine GEMM split over the ranks, all-gather of the score slices',
                'trials': 1 << 20, 'seconds': dt4, 'trials_per_s': (1 << 20) / dt4, 'embeddings_per_s': M4 / dt4,
                'score_checksum': float(scores4.double().sum().item()), 'timing': 'CUDA events, max over ranks, best of 2'}
        del x4

    # ---- DoubleMHA pooling microbench (BASELINE configs[1]): B=512, T=200, D=1024, H=16, length-masked
    dmha = None
    if not args.no_dmha and rank == 0:
        dmha = {}
        Bp, Tp, Dp, Hp = 512, 200, 1024, 16
        gen = torch.Generator(device=dev).manual_seed(0)
        q = torch.randn(Dp // Hp, Hp, device=dev, generator=gen) * 0.3
        a = torch.randn(Dp // Hp, device=dev, generator=gen) * 0.3
        lens = torch.from_numpy(synth.make_lengths(Bp, 100, 200, seed=0)).to(dev)
        def time_us(fn, reps):
            """Average device time of fn(i): replayed from a CUDA graph so the Python/ctypes launch path (~100 us per
            call, comparable to these kernels) is not in the measurement; eager timing if capture is unavailable."""
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(reps):
                        fn(i)
                g.replay()
                torch.cuda.synchronize()
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                mode = 'cuda-graph'
            except Exception:
                torch.cuda.synchronize()
                e0.record()
                for i in range(reps):
                    fn(i)
                e1.record()
                torch.cuda.synchronize()
                mode = 'eager'
            return e0.elapsed_time(e1) * 1e3 / reps, mode

        for dt_name, dt in (('fp32', torch.float32), ('bf16', torch.bfloat16)):
            bufs = [torch.randn(Bp, Tp, Dp, device=dev, generator=gen).to(dt) for _ in range(2)]   # 2 x 419 MB fp32 >> L2
            es = 4 if dt == torch.float32 else 2
            lens_sorted = torch.sort(lens, descending=True).values   # longest first: the order extract.bucket_plan builds batches in
            for case, L in (('full', None), ('masked', lens), ('masked_sorted', lens_sorted)):
                nbytes = (Bp * Tp if L is None else int(L.sum().item())) * Dp * es + Bp * (Dp // Hp) * 4
                us, mode = time_us(lambda i: ops.dmha_fwd(bufs[i & 1], q, a, lengths=L, need_align=False), 20)
                gbs = nbytes / (us * 1e-6) / 1e9
                dmha['%s_%s' % (dt_name, case)] = {'us': us, 'gbs': gbs, 'frac': gbs / pk['hbm_gbs'], 'bytes': nbytes, 'timing': mode}
            # backward (read x, write dx): 2x the forward's bytes
            r = ops.dmha_fwd(bufs[0], q, a, need_align=False)
            gout = torch.randn(Bp, Dp // Hp, device=dev, generator=gen)
            us, mode = time_us(lambda i: ops.dmha_bwd(bufs[i & 1], q, a, gout, None, r['ctx'], r['lse'], r['headw']), 10)
            nbytes = 2 * Bp * Tp * Dp * es
            dmha['%s_bwd' % dt_name] = {'us': us, 'gbs': nbytes / (us * 1e-6) / 1e9, 'frac': nbytes / (us * 1e-6) / 1e9 / pk['hbm_gbs'],
                                        'bytes': nbytes, 'timing': mode, 'note': 'includes the dquery/datt reduce kernel'}
            del bufs
        dmha['peak_gbs'] = pk['hbm_gbs']
        dmha['shape'] = 'B=512 T=200 D=1024 H=16; 2 rotating inputs (each > L2); no alignment output'

    # ---- secondary measurements (not the headline): the rows SURVEY.md §8(f) marks "next"
    extras = None
    if not args.no_extras and not args.no_dmha and rank == 0:
        extras = {}
        try:
            from doubleattentionspeakerverification_b200 import CNNs, poolings, featureExtractor as fe
            torch.cuda.empty_cache()
            tb = 128
            tnet = CNNs.VGG4L(1024, precision='bf16', train_kernels=True).to(dev)
            tpool = poolings.DoubleMHA(5120, 32, mask_prob=0.3).to(dev).train()
            tx = torch.randn(tb, FRAMES, 80, device=dev) * 2

            def train_step():
                tnet.zero_grad(set_to_none=True); tpool.zero_grad(set_to_none=True)
                o, _ = tpool(tnet(tx))
                o.square().mean().backward()
            for _ in range(2):
                train_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                train_step()
            e1.record(); torch.cuda.synchronize()
            tms = e0.elapsed_time(e1) / 5
            extras['train_step'] = {'what': 'VGG4L(1024) + DoubleMHA forward + backward on the package kernels (train_kernels=True), synthetic',
                                    'batch': tb, 'ms': tms, 'utterances_per_s': tb / tms * 1e3,
                                    'conv_tflops': 3.0 * sum(conv_flops(tb, FRAMES).values()) / (tms * 1e-3) / 1e12}
            del tnet, tpool, tx
            torch.cuda.empty_cache()
            n = 512 + 160 * (FRAMES - 1)
            wave = torch.from_numpy(np.stack([synth.make_waveform(n, 16000, seed=i) for i in range(4)] * 64).astype(np.float32)).to(dev)
            for _ in range(3):
                fe.logmel_batch(wave, [n] * 256, 16000)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fe.logmel_batch(wave, [n] * 256, 16000)
            e1.record(); torch.cuda.synchronize()
            fms = e0.elapsed_time(e1) / 10
            extras['logmel'] = {'what': '256 waveforms of %d samples (16 kHz) -> [256,%d,80] log-mel + CMN on the GPU' % (n, FRAMES),
                                'ms': fms, 'utterances_per_s': 256 / fms * 1e3}
        except Exception as e:                                   # secondary numbers must never take the headline down
            extras['error'] = repr(e)[:300]

    # ---- the other precisions of the same step (fp16 operands; fp32 parity path) as secondary numbers
    if extras is not None and 'error' not in extras:
        try:
            for prec, bsz, reps in (('fp16', Bn, 5), ('fp32x3', Bn, 3), ('fp32', 32, 2)):
                if prec == args.precision:
                    continue
                c2 = synth.example_config()
                c2.precision = prec
                n2 = synth.load_state_dict(model.SpeakerClassifier(c2, dev), synth.make_state_dict(c2, 1234)).to(dev).eval()
                xb = xs[0][:bsz].contiguous()
                with torch.no_grad():
                    for _ in range(3):
                        n2.getEmbedding(xb)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        n2.getEmbedding(xb)
                    e1.record()
                    torch.cuda.synchronize()
                t2 = e0.elapsed_time(e1) / reps
                extras['precision_' + prec] = {'batch': bsz, 'ms_per_step': t2, 'embeddings_per_s': bsz / t2 * 1e3,
                                               'what': 'the same getEmbedding step with precision=%r' % prec}
                del n2
                torch.cuda.empty_cache()
        except Exception as e:
            extras['precision_error'] = repr(e)[:300]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, n, el, threads = cpu_reference_rate(15.0)
        cpu = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': '%d utterances (400x80) in %.1f s through oracle/torch_port.get_embedding, fp32' % (n, el)}

    if rank == 0:
        h2d = x_host.numel() * 4
        d2h = emb_host.numel() * 4
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': {'bf16': 'bf16', 'fp16': 'f16', 'fp32': 'f32', 'fp32x3': 'f32 (3 x bf16 split)'}[args.precision], 'data': 'synthetic',
                'config': workload_config(world, Bn),
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'ms_per_step': ms_e2e / args.steps},
                'gpu_launches': launches, 'clocks': clk.summary(), 'roofline': roof, 'dmha_microbench': dmha, 'configs3': cfg3, 'configs4': cfg4, 'extras': extras, 'cpu_baseline': cpu}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
