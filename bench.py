#!/usr/bin/env python
"""bench.py -- utterances/s of the STFT -> mask -> PIT hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]
                    [--config pit|pit3|dc|tasnet|train]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config selects the workload: `pit` (default) is BASELINE.json's headline, described below; the others are
BASELINE.json's configs 2-5 (bench_configs.py): `pit3` K = 3 / 8 s, `dc` deep-clustering loss on a ragged batch,
`tasnet` time-domain PIT losses + the gradient all-reduce of a ConvTasNet, `train` one full training step of the
BLSTM mask estimator (cuDNN LSTM + tcgen05 projections + fused loss + overlapped NCCL gradient exchange).
--impl reference-gpu times the unmodified reference ops on the same GPU (its ATen composition).

One "step" = one pass of the hot path over one batch of synthetic 4 s / 16 kHz 2-speaker mixtures:
  kernel 1  |Y| = |STFT(y)|                      (front-end feature, b2s_stft_forward, ABS epilogue)
  kernel 2  fused STFT(s_k) -> mask (*) |Y| -> K x K SSE -> permutation search (b2s_stft_pit_forward)
The mask network between the two is NOT part of the path: masks are
synthetic U(0,1) tensors resident in HBM.  Rank r of N processes its own batch (weak scaling, no
data-path collective: utterances are independent, DESIGN.md "Multi-GPU").

Prints ONE JSON line (see the task contract): value = device-timed utterances/s with inputs resident
in HBM; e2e = the same through the public call with pinned HOST buffers, H2D/D2H inside the timed
region; roofline = achieved algorithmic GB/s of the dominant kernel vs the measured HBM peak;
cpu_baseline = the CPU oracle port of the reference ops timed on this box's host cores.
--impl reference times that CPU implementation as the headline instead.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL reads its environment when the library is loaded (import torch): ask for the INFO log (ranks, rings / trees /
# NVLS) before that; main() points file descriptor 1 at stderr, so the log never mixes with the JSON line on stdout
# (the GPU boxes export NCCL_DEBUG=VERSION: override it, B2S_NCCL_DEBUG selects another level)
os.environ['NCCL_DEBUG'] = os.environ.get('B2S_NCCL_DEBUG', 'INFO')

import numpy as np  # noqa: E402
import torch  # noqa: E402

BATCH = 64            # utterances per GPU per step (north star: batch 64 x 4 s x 16 kHz)
SAMPLES = 64000       # 4 s @ 16 kHz
SOURCES = 2
SIZE, SHIFT = 1024, 256
FRAMES, BINS = 253, 513
ROTATE = 3            # input sets cycled so a step's 215 MB of inputs were evicted from the 126 MB L2
GRAPH_STEPS = 12      # steps captured in one CUDA graph (a multiple of ROTATE)
WORKLOAD = ('fused STFT->mask->PIT-loss path, batch 64 x 4 s x 16 kHz, 2 speakers, STFT(1024,256) '
            '(253 frames x 513 bins); masks synthetic U(0,1), mask network excluded (SURVEY 8d(i))')

# algorithmic (compulsory) bytes per utterance, SURVEY.md section 8(d) / BASELINE.md section 2
BYTES_FRONT = 4 * SAMPLES + 4 * FRAMES * BINS                                  # y -> |Y|
BYTES_LOSS = 4 * SAMPLES * (1 + SOURCES) + 4 * FRAMES * BINS * SOURCES         # mask, y, s -> loss, perm
BYTES_PATH = BYTES_FRONT + BYTES_LOSS                                          # 2 581 468 B/utt
FUSED_WARP_INSTRUCTIONS = 23.49e6   # smsp__inst_executed.sum of one fused launch at this shape (ncu, profiles/)


def headline_config(world):
    """Identical in every arm (ours / reference / reference-gpu): the workload, not how an arm runs it."""
    return {'workload': WORKLOAD, 'batch_per_gpu': BATCH, 'samples': SAMPLES, 'sources': SOURCES,
            'frames': FRAMES, 'bins': BINS,
            'l2': f'inputs larger than L2: {ROTATE} rotating input sets of 215 MB each',
            'parallelism': f'{world} independent shard(s), no data-path collective'}


_JSON_OUT = None


def reserve_stdout_for_json():
    """From here on file descriptor 1 points at stderr: whatever libraries print (NCCL's INFO log goes to stdout)
    lands in stderr, and emit() writes the one JSON line to the ORIGINAL stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit(line):
    text = json.dumps(line)
    if _JSON_OUT is not None:
        _JSON_OUT.write(text + '\n')
        _JSON_OUT.flush()
    else:
        print(text, flush=True)


def nccl_info_to_stderr():
    """NCCL's INFO log (ranks, rings / trees / NVLS) is kept, not silenced (NCCL_DEBUG is set at the top of this file,
    before torch is imported); it reaches stderr through reserve_stdout_for_json()."""
    assert _JSON_OUT is not None, 'call reserve_stdout_for_json() first'


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as fd:
            return float(json.load(fd)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    BAD = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown')

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.thread, self.nvml, self.handle, self.max_mhz = None, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None

    def _reasons(self):
        n = self.nvml
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            try:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            except Exception:
                return []
        table = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
                 'sw_power_cap': 0x4, 'hw_power_brake': 0x80}
        return [name for name, bit in table.items() if mask & bit]

    def _run(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                mhz = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                self.samples.append((time.perf_counter(), mhz, self._reasons()))
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread is not None:
            self.thread.join()

    def summary(self, t0, t1):
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not inside:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': [], 'samples': 0}
        reasons = sorted({r for s in inside for r in s[2]})
        return {'sm_mhz': statistics.median(s[1] for s in inside), 'sm_max_mhz': self.max_mhz,
                'reasons': reasons, 'samples': len(inside)}


# ------------------------------------------------------------------------------------------------ data
def synthetic_batch(seed, device=None, pin=False):
    """s ~ 0.1 N(0,1) [B, K, T], y = sum_k s_k, masks ~ U(0,1) [B, M, K, F]."""
    gen = torch.Generator().manual_seed(seed)
    s = 0.1 * torch.randn(BATCH, SOURCES, SAMPLES, generator=gen)
    y = s.sum(1)
    masks = torch.rand(BATCH, FRAMES, SOURCES, BINS, generator=gen)
    out = dict(y=y, s=s, masks=masks)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out


# ------------------------------------------------------------------------------------------------ CPU arm
def import_reference():
    """The unmodified reference package if it was pip-installed into baseline/_ref, with the
    stand-ins of oracle/ref_standins for its un-vendored dependencies; None otherwise."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'padertorch')):
        return None
    for path in (os.path.join(ROOT, 'oracle', 'ref_standins'), ref):
        if path not in sys.path:
            sys.path.insert(0, path)
    try:
        import warnings
        warnings.filterwarnings('ignore')
        import padertorch
        return padertorch
    except Exception:
        return None


def reference_step(pt, batch, stft, n_utt):
    """The reference's own code for the step: pt.ops.STFT on y and s, abs (pit/data.py:49-77), then the
    per-example pit_loss loop of pit/model.py:117-128."""
    with torch.no_grad():
        y_abs = stft(batch['y'][:n_utt]).abs()
        x_abs = stft(batch['s'][:n_utt]).abs().transpose(1, 2)
        out = []
        for b in range(n_utt):
            out.append(pt.ops.losses.pit_loss(batch['masks'][b] * y_abs[b][:, None, :], x_abs[b], axis=-2,
                                              return_permutation=True))
        return out


def cpu_step(batch, stft, n_utt):
    """Oracle port of the same step (used when the reference package is not installed)."""
    from oracle import path as oracle_path
    with torch.no_grad():
        return oracle_path.stft_mask_pit_step(batch['y'][:n_utt], batch['s'][:n_utt],
                                              batch['masks'][:n_utt], stft=stft)


def time_cpu(steps, warmup, n_utt=BATCH, seed=1234):
    """Returns (utt/s, seconds per step, threads, kind) of the CPU implementation of the step."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = synthetic_batch(seed)
    pt = import_reference()
    if pt is not None:
        stft = pt.ops.STFT(SIZE, SHIFT)
        run, kind = (lambda: reference_step(pt, batch, stft, n_utt)), 'reference'
    else:
        from oracle.stft import ReferenceSTFT
        stft = ReferenceSTFT(SIZE, SHIFT)
        run, kind = (lambda: cpu_step(batch, stft, n_utt)), 'port'
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    elapsed = time.perf_counter() - t0
    return n_utt * steps / elapsed, elapsed / steps, cores, kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    value, per_step, cores, kind = time_cpu(args.steps, args.warmup)
    cpu_model = ''
    try:
        with open('/proc/cpuinfo') as fd:
            cpu_model = next(l.split(':', 1)[1].strip() for l in fd if l.startswith('model name'))
    except Exception:
        pass
    line = {
        'impl': 'reference', 'metric': 'utterances/sec', 'value': value, 'unit': 'utt/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': per_step * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': headline_config(world),
        'cpu_baseline': {'value': value, 'unit': 'utt/s', 'cores': cores, 'kind': kind,
                         'sample': f'{args.steps} x one full batch of {BATCH} utterances through '
                                   + ('the unmodified reference (baseline/_ref): ' if kind == 'reference'
                                      else 'the oracle port of ')
                                   + f'pt.ops.STFT + per-example pit_loss loop, {cores} threads, {cpu_model}'},
        'e2e': {'value': value, 'unit': 'utt/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def run_reference_gpu(args, rank, world, local_rank):
    """The reference's own ATen composition on the B200: the unmodified
    pt.ops.STFT (dense DFT convolution) + per-example pit_loss loop on CUDA tensors, device-timed."""
    if rank != 0:
        return
    pt = import_reference()
    if pt is None:
        emit({'impl': 'reference-gpu', 'unavailable': 'reference not installed in baseline/_ref'})
        return
    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    stft = pt.ops.STFT(SIZE, SHIFT)
    sets = [synthetic_batch(i, device=device) for i in range(ROTATE)]
    for i in range(max(args.warmup, 3)):
        reference_step(pt, sets[i % ROTATE], stft, BATCH)
    torch.cuda.synchronize()
    steps = max(3, min(args.steps, 20))
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(steps):
        reference_step(pt, sets[i % ROTATE], stft, BATCH)
    end.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(end) / steps
    value = BATCH / (ms * 1e-3)
    emit({'impl': 'reference-gpu', 'metric': 'utterances/sec', 'value': value, 'unit': 'utt/s', 'n_gpus': 1,
                      'steps': steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True,
                      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': headline_config(1),
                      'note': 'unmodified reference ops (pt.ops.STFT = conv1d with the dense DFT matrix, per-example '
                              'pit_loss loop with int(idx) syncs) on CUDA tensors, python eager as the reference runs them',
                      'gpu_launches': 0})


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import padertorch_b200 as b2s
    from padertorch_b200 import review

    device = torch.device('cuda', local_rank)
    torch.cuda.set_device(device)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        nccl_info_to_stderr()
        dist.init_process_group('nccl', device_id=device)

    stft = b2s.ops.STFT(SIZE, SHIFT)
    sets = [synthetic_batch(100 * rank + i, device=device) for i in range(ROTATE)]
    torch.cuda.synchronize()

    def step(data):
        y_abs = stft.magnitude(data['y'])                                        # kernel 1
        return review.stft_mask_pit_step(None, data['s'], data['masks'], stft=stft,
                                         observation_abs=y_abs)                  # kernel 2

    for i in range(max(args.warmup, 3)):
        loss, perm = step(sets[i % ROTATE])
    torch.cuda.synchronize()

    # The two launches of a step cost ~10x more host time through Python than the kernels run on the
    # device, so the steady-state loop replays one CUDA graph per input set (plans, workspaces and
    # output buffers are created by the warm-up above / inside the capture pool).
    # One more graph holds GRAPH_STEPS consecutive steps (whole rounds of the input sets): inside it the front-end
    # of step n + 1 is chained to the fused kernel of step n by programmatic dependent launch like the two kernels
    # of a step are (its launch and prologue overlap the predecessor's tail); K steps = K // GRAPH_STEPS replays
    # of it + K % GRAPH_STEPS single-step graphs.
    graphs, results, round_graph = [], [], None
    if not args.eager:
        for data in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                results.append(step(data))
            graphs.append(g)
        if not args.single_stream and not args.single_step_graphs:
            # The front-end of a step does not depend on the loss of the previous step (the features of batch n + 1
            # are prepared while the loss of batch n is computed): its kernels are captured on a second stream,
            # ordered only by their true dependencies, and fill the ragged tail of the running fused kernel.
            round_graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device)
            with torch.cuda.graph(round_graph):
                main = torch.cuda.current_stream()
                fork = torch.cuda.Event()
                fork.record(main)
                side.wait_event(fork)
                # every step keeps its own |Y| buffer alive for the life of the graph: a buffer freed during the
                # capture could be handed to the next front-end on the side stream while the loss kernel on the
                # main stream still reads it
                round_results, round_features = [], []
                for j in range(GRAPH_STEPS):
                    data = sets[j % ROTATE]
                    with torch.cuda.stream(side):
                        y_abs = stft.magnitude(data['y'])
                        ready = torch.cuda.Event()
                        ready.record(side)
                    round_features.append(y_abs)
                    main.wait_event(ready)
                    round_results.append(review.stft_mask_pit_step(None, data['s'], data['masks'], stft=stft,
                                                                   observation_abs=y_abs))
        elif not args.single_step_graphs:
            round_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(round_graph):
                round_results = [step(sets[j % ROTATE]) for j in range(GRAPH_STEPS)]
        torch.cuda.synchronize()

    def run_step(i):
        if graphs:
            graphs[i % ROTATE].replay()
            return results[i % ROTATE]
        return step(sets[i % ROTATE])

    def run_steps(n):
        """Exactly n steps, cycling through the input sets; returns the last step's result."""
        out, i = None, 0
        if round_graph is not None:
            for _ in range(n // GRAPH_STEPS):
                round_graph.replay()
            i = n - n % GRAPH_STEPS
            if i:
                out = round_results[-1]
        for j in range(i, n):
            out = run_step(j)
        return out

    loss, perm = run_steps(max(args.warmup, 3))
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    start.record()
    loss, perm = run_steps(args.steps)
    end.record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    elapsed_ms = start.elapsed_time(end)
    if distributed:
        t = torch.tensor([elapsed_ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
        dist.barrier()

    # the captured multi-step graph must reproduce the eagerly launched step bit for bit (not part of any number)
    if round_graph is not None:
        round_graph.replay()
        torch.cuda.synchronize()
        for j in range(GRAPH_STEPS):
            want_loss, want_perm = step(sets[j % ROTATE])
            got_loss, got_perm = round_results[j]
            assert torch.equal(got_loss, want_loss) and torch.equal(got_perm, want_perm), 'captured graph differs from the eager step'

    # keep the GPU under the same load long enough for NVML to see it (not part of any number)
    t_hold = time.perf_counter()
    i = 0
    while time.perf_counter() - t_hold < 1.0:
        run_steps(63)
        torch.cuda.synchronize()
    torch.cuda.synchronize()
    t_wall2 = time.perf_counter()
    sampler.stop()
    clocks = sampler.summary(t_wall0, t_wall2)
    clocks['window'] = 'timed region + 1 s of identical steps'

    # ---- per-kernel durations (CUDA events on the launching stream), for the roofline object.
    # Each kernel is replayed from its own one-node CUDA graph so that Python launch overhead does not
    # sit between the two events; inputs rotate exactly as in the timed loop.
    def kernel_ms(fn):
        gs = []
        for data in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn(data)
            gs.append(g)
        for i in range(6):
            gs[i % ROTATE].replay()
        # CUDA event timestamps tick at ~2 us on this platform: time bursts of replays, not single ones
        burst = 4 * ROTATE
        n_probe = max(5, min(max(args.steps, 30), 300) // burst)
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                 for _ in range(n_probe)]
        for e0, e1 in pairs:
            e0.record()
            for i in range(burst):
                gs[i % ROTATE].replay()
            e1.record()
        torch.cuda.synchronize()
        return statistics.median(e0.elapsed_time(e1) for e0, e1 in pairs) / burst

    y_abs_sets = [stft.magnitude(d['y']) for d in sets]
    front_ms = kernel_ms(lambda d: stft.magnitude(d['y']))
    idx = {id(d): i for i, d in enumerate(sets)}
    loss_ms = kernel_ms(lambda d: review.stft_mask_pit_step(None, d['s'], d['masks'], stft=stft,
                                                            observation_abs=y_abs_sets[idx[id(d)]]))

    # ---- end to end through the public call with HOST buffers
    host_sets = [synthetic_batch(100 * rank + i, pin=True) for i in range(2)]
    h2d = sum(v.numel() * 4 for v in host_sets[0].values())
    out_loss = torch.empty(BATCH, dtype=torch.float32).pin_memory()
    out_perm = torch.empty((BATCH, SOURCES), dtype=torch.int32).pin_memory()
    d2h = out_loss.numel() * 4 + out_perm.numel() * 4

    # double buffered: the host -> device copies of step i + 1 run on a copy stream while step i computes; every step
    # still moves its own inputs over PCIe and reads its own results back inside the timed region
    copy_stream = torch.cuda.Stream(device)
    slots = [dict(data={k: torch.empty_like(v, device=device) for k, v in host_sets[0].items()},
                  ready=torch.cuda.Event(), free=torch.cuda.Event()) for _ in range(2)]

    def upload(i):
        slot = slots[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot['free'])          # the step that last used this slot is done with it
            for k, v in host_sets[i % 2].items():
                slot['data'][k].copy_(v, non_blocking=True)
            slot['ready'].record(copy_stream)

    def e2e_loop(n):
        for slot in slots:
            slot['free'].record(torch.cuda.current_stream())
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            slot = slots[i % 2]
            torch.cuda.current_stream().wait_event(slot['ready'])
            loss, perm = step(slot['data'])
            out_loss.copy_(loss, non_blocking=True)
            out_perm.copy_(perm, non_blocking=True)
            slot['free'].record(torch.cuda.current_stream())

    e2e_steps = max(5, min(args.steps, 30))
    e2e_loop(3)
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if distributed:
        t = torch.tensor([e2e_s], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())

    # The headline path has no data-path collective (utterances are independent).  What a data-parallel TRAINING step on
    # it exchanges is the mask network's gradient: measured here, outside every timed region, as the bucketed sum
    # all-reduce of the PIT BLSTM model's 101.3 MB of gradients (padertorch_b200.parallel's exchange; --config train
    # measures it inside a real training step).
    probe = None
    if distributed:
        nbytes, bucket = 101298508, 32 << 20
        flat = torch.zeros(nbytes // 4, dtype=torch.float32, device=device)
        chunks = list(flat.split(bucket // 4))
        for _ in range(3):
            for c in chunks:
                dist.all_reduce(c)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            for c in chunks:
                dist.all_reduce(c)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        probe = {'op': 'all_reduce(sum), fp32, 32 MiB buckets', 'bytes': nbytes, 'ranks': world, 'ms': ms,
                 'algbw_gbs': nbytes / (ms * 1e-3) / 1e9, 'busbw_gbs': nbytes / (ms * 1e-3) / 1e9 * 2 * (world - 1) / world,
                 'note': 'gradient exchange of the PIT BLSTM mask estimator (25.3 M parameters); device-timed, max over '
                         'ranks, outside the timed region of `value` and `e2e`'}
        del flat, chunks

    if rank == 0:
        peak, peak_kind = measured_peaks()
        ms_per_step = elapsed_ms / args.steps
        value = world * BATCH * args.steps / (elapsed_ms * 1e-3)
        achieved = BYTES_LOSS * BATCH / (loss_ms * 1e-3) / 1e9
        path_gbs = BYTES_PATH * BATCH / (ms_per_step * 1e-3) / 1e9
        cpu_value, cpu_step_s, cores, cpu_kind = time_cpu(3, 1)
        line = {
            'metric': 'utterances/sec', 'value': value, 'unit': 'utt/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms_per_step,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': headline_config(world),
            'launch': 'python eager' if args.eager else ('CUDA graph replay (2 kernel nodes per step)' if args.single_step_graphs else f'CUDA graph replay ({GRAPH_STEPS} steps = {2 * GRAPH_STEPS} kernel nodes per graph, 2 kernel nodes per step' + ('' if args.single_stream else '; front-end kernels captured on a second stream: front-end of step n + 1 overlaps the tail of the loss kernel of step n') + ')'),
            'e2e': {'value': world * BATCH * e2e_steps / e2e_s, 'unit': 'utt/s', 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h, 'steps': e2e_steps},
            'gpu_launches': 2 * args.steps,
            'clocks': clocks,
            'roofline': {'bound': 'hbm', 'kernel': 'stft_pit_pair_kernel', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None, 'peak_kind': peak_kind,
                         'algorithmic_bytes_per_launch': BYTES_LOSS * BATCH, 'kernel_ms': loss_ms},
            'roofline_issue': {'bound': 'fp32-issue', 'kernel': 'stft_pit_pair_kernel',
                               'achieved': FUSED_WARP_INSTRUCTIONS / (loss_ms * 1e-3) / 1e9,
                               'peak': 148 * 4 * 1.965, 'unit': 'G warp-instructions/s',
                               'frac': FUSED_WARP_INSTRUCTIONS / (loss_ms * 1e-3) / 1e9 / (148 * 4 * 1.965),
                               'note': 'issue slots: 148 SMs x 4 sub-partitions x 1.965 GHz; instructions per launch '
                                       'from ncu smsp__inst_executed.sum (profiles/r2_ncu_summary.txt); the kernel is '
                                       'bound by shared-memory + FP32-pipe work per frame, not by HBM (DESIGN.md 6)'},
            'path_roofline': {'achieved': path_gbs, 'frac': path_gbs / peak, 'unit': 'GB/s',
                              'algorithmic_bytes_per_step': BYTES_PATH * BATCH,
                              'front_end_kernel_ms': front_ms,
                              'front_end_frac': BYTES_FRONT * BATCH / (front_ms * 1e-3) / 1e9 / peak},
            'cpu_baseline': {'value': cpu_value, 'unit': 'utt/s', 'cores': cores, 'kind': cpu_kind,
                             'sample': f'3 x one full batch of {BATCH} utterances ('
                                       + ('unmodified reference from baseline/_ref: ' if cpu_kind == 'reference'
                                          else 'oracle port of ')
                                       + f'pt.ops.STFT + per-example pit_loss loop), {cpu_step_s:.2f} s per batch'},
        }
        traffic_file = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(traffic_file):
            with open(traffic_file) as fd:
                line['roofline']['traffic'] = json.load(fd).get('stft_pit_pair_kernel')
        if probe is not None:
            line['collective_probe'] = probe
        emit(line)
    if distributed:
        dist.destroy_process_group()


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=1000)
    parser.add_argument('--warmup', type=int, default=10)
    parser.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference-gpu'])
    parser.add_argument('--config', default='pit', choices=['pit', 'pit3', 'dc', 'tasnet', 'train'],
                        help="workload: BASELINE.json's headline (pit) or one of its configs 2-5")
    parser.add_argument('--eager', action='store_true', help='launch from Python instead of CUDA graphs')
    parser.add_argument('--single-stream', action='store_true', help='capture the multi-step graph on one stream (no overlap of the next front-end with the running loss kernel)')
    parser.add_argument('--single-step-graphs', action='store_true', help='one CUDA graph per step instead of one per round of input sets')
    args = parser.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    reserve_stdout_for_json()
    if args.config != 'pit':
        sys.modules.setdefault('bench', sys.modules[__name__])   # bench_configs shares this module's state (emit)
        import bench_configs
        bench_configs.run(args, rank, world, local_rank)
        return
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: padertorch_b200 has no CPU fallback')
    if args.impl == 'reference-gpu':
        run_reference_gpu(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
