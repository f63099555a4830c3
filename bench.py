"""bench.py -- captions/sec of the GIT captioning hot path (BASELINE.json metric).

A "step" is one batch of 64 synthetic 224x224 images captioned through the reference surface: CLIP-ViT encoder ->
visual projection -> image-row prefill of the 6 decoder layers -> 39 KV-cached greedy decode steps
(max_len 40), i.e. the reference's `CaptioningModel.forward` in eval mode with its greedy decoder
(reference model.py:27-33).  N=1 workload = BASELINE.json configs[1]: GIT_BASE, batch 64, one B200.
Random-init weights of that architecture (reference initialiser distributions) and synthetic pixels.

Three numbers per run: `sync_value` = `model(batch)` one batch at a time (the reference's calling pattern);
`value` = the same batches handed to `model.submit(batch, depth, coalesce)` (device-resident pixels): `coalesce` batches
share one engine launch, `depth` launches are in flight; `e2e` = the same with pinned HOST tensors in and tokens read back.
Every batch's full work (encoder, prefill, 39 decode steps, search) is inside the timed region in all three.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Multi-GPU (torchrun, one rank per GPU): every rank captions its own batch (weak scaling, image-wise
sharding, reference inference.py:165-169) and the timed region ends with ONE NCCL all_gather of the
finished token ids.  `--impl reference` times the reference's own CPU algorithm (the as-shipped, no-KV-cache
restatement in oracle/git_oracle.py -- the Python reference itself cannot travel to the GPU box) on the
host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'captions/sec (greedy, max_len=40) GIT_BASE batch64'
UNIT = 'captions/s'
MAX_STEPS = 40
BATCH = 64


class Tok:
    cls_token_id, sep_token_id = 101, 102


def env_int(name, default):
    return int(os.environ.get(name, default))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], bf16_tflops=d['bf16_tflops'], bf16_sustained=d.get('bf16_tflops_sustained'),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source='fallback (B200_PROFILING.md)')


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (per launch) from the committed
    `ncu --set full` capture (profiles/prof_gemm_r01.md): 24.1 MB read + 24.8 MB write."""
    p = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if os.path.exists(p):
        return json.load(open(p)).get('dram_bytes_per_launch')
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line), streamed with
    `-lms` so that even a sub-second region gets several samples."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.proc = None
        self.stop_flag = False

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '20'], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                f = [x.strip() for x in line.strip().split(',')]
                if len(f) < 6:
                    continue
                try:
                    self.samples.append(float(f[0]))
                    self.max_mhz = float(f[1])
                except ValueError:
                    continue
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith('active'):
                        self.reasons.add(n)
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        s = sorted(self.samples)
        return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(s)}


def cpu_baseline_run(sample_batch, steps, warmup, threads=None):
    """The reference's CPU path as shipped (full [image || text] recompute every step), fp32, all host threads."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import git_oracle
    from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images
    threads = threads or min(os.cpu_count(), 16)
    torch.set_num_threads(threads)
    sd = synthetic_state_dict({}, 0, 'init')
    img = synthetic_images(sample_batch, 0, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        out = git_oracle.generate(sd, {}, {'image': img}, 'greedy', MAX_STEPS, cached=False)
        dt = time.perf_counter() - t0
        assert out['predictions'].shape == (sample_batch, MAX_STEPS)
        if i >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    return sample_batch / mean, mean, threads


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    sample = 4
    value, sec, threads = cpu_baseline_run(sample, max(1, min(args.steps, 3)), 1 if args.warmup > 0 else 0)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'GIT_BASE greedy max_len=40, synthetic 224x224, random-init weights', 'global_batch': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': 'oracle/git_oracle.py as-shipped mode (no KV cache, fp32 torch CPU ops): batch of %d images '
                                   'per step instead of 64 (same per-image work; CPU throughput is batch-insensitive here)' % sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def main():
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get('GITB200_BENCH_WATCHDOG_S', '480')), exit=True)   # a hung run reports where
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=16)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='gitb200')
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-micro', action='store_true')
    ap.add_argument('--pipeline', type=int, default=2, choices=[1, 2, 3, 4, 5, 6, 7, 8],
                    help='engine launches in flight (the encoder of launch i+1 overlaps the decode loop of launch i)')
    ap.add_argument('--coalesce', type=int, default=4, choices=[1, 2, 3, 4],
                    help='dynamic batching: this many submitted batches of 64 share one engine launch (one decode chain over '
                         'all their rows). Measured on B200 (tools/batch_sweep.py, profiles/batch_sweep_r01.txt): 64 x 4 in '
                         'flight 5319 captions/s, 128 x 3 6086, 256 x 2 6353, 256 x 3 6432')
    ap.add_argument('--ncu-range', action='store_true',
                    help='bracket the timed region with cudaProfilerStart/Stop (use with ncu --profile-from-start off)')
    args = ap.parse_args()
    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)
    args.coalesce = max(1, min(args.coalesce, 256 // max(1, args.batch)))    # one decode chain serves at most 256 rows

    import ctypes
    import torch
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.build()
    from generativeimage2text_b200 import _lib
    from generativeimage2text_b200.model import get_git_model, AutoRegressiveBeamSearch
    from generativeimage2text_b200.sharding import gather_captions
    from generativeimage2text_b200.synthetic import synthetic_state_dict, synthetic_images

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    B = args.batch
    os.environ.setdefault('GITB200_SLOTS', str(max(4, args.pipeline)))
    model = get_git_model(Tok(), {})
    model.load_state_dict(synthetic_state_dict({}, 0, 'init'), strict=True)
    model = model.to(dev).eval()
    model.decoder = AutoRegressiveBeamSearch(102, max_steps=MAX_STEPS, beam_size=1, per_node_beam_size=1,
                                             fix_missing_prefix=True)
    img_host = synthetic_images(B, 0, 1234 + rank).contiguous().pin_memory()
    img_dev = img_host.to(dev)
    stream = torch.cuda.Stream(device=dev)
    n_total = B * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run_device(k, depth, src=None, to_host=False):
        """k steps; depth 1 = one batch at a time (model(batch)); depth > 1 = up to `depth` engine launches in flight, each
        serving `coalesce` submitted batches, so the encoder of later batches overlaps the latency-bound decode loops of
        earlier ones.  src: the step's input (device-resident pixels, or the pinned host tensor for the e2e leg)."""
        src = img_dev if src is None else src
        pend, toks = [], None

        def collect(p):
            out = p.result()
            t, l = out['predictions'], out['logprobs']
            if to_host:                      # the caller reads the result: D2H inside the timed region
                t, l = t.cpu(), l.cpu()
                if world > 1:
                    t, l = t.to(dev), l.to(dev)
            if world > 1:
                t, l = gather_captions(t, l, n_total)
            return t
        for _ in range(k):
            if depth == 1:     # the reference call: model(batch), one at a time on the caller's stream
                out = model({'image': src})
                toks = out['predictions']
                if world > 1:
                    toks, _ = gather_captions(toks, out['logprobs'], n_total)
                continue
            pend.append(model.submit({'image': src}, depth=depth, coalesce=args.coalesce))
            if len(pend) >= depth * args.coalesce:
                toks = collect(pend.pop(0))
        while pend:
            toks = collect(pend.pop(0))
        return toks

    # ---------------- device-resident timing (`value`) ----------------
    results = {}
    with torch.cuda.stream(stream):
        for depth in (1, args.pipeline):
            if depth in results:
                continue
            toks = run_device(max(args.warmup, 2 * depth * args.coalesce), depth)   # every engine slot past its first (capturing) call
            tail = args.steps % args.coalesce
            if depth > 1 and args.coalesce > 1 and tail:
                # the last launch of the timed region serves only `tail` batches: let every slot capture the decode-step
                # graph of that row count too (warm-up, like the full-size launches above)
                small = torch.cat([img_dev] * tail, dim=0)
                for k in range(depth):
                    model.submit({'image': small}, slot=k).result()
                del small
            barrier()
            if depth == args.pipeline:
                sampler = ClockSampler(local)
                if rank == 0:
                    sampler.start()
            launches0 = model.launch_count()
            if args.ncu_range and depth == args.pipeline:
                torch.cuda.profiler.start()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            toks = run_device(args.steps, depth)
            torch.cuda.current_stream().wait_stream(stream)
            for sl in model._slots:          # the pipelined engines run on their own streams: join them before e1
                if sl['stream'] is not None:
                    stream.wait_stream(sl['stream'])
            e1.record(stream)
            barrier()
            if args.ncu_range and depth == args.pipeline:
                torch.cuda.profiler.stop()
            ms_d = e0.elapsed_time(e1)
            t = torch.tensor([ms_d], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            results[depth] = (t.item(), model.launch_count() - launches0)
            assert toks.shape[0] == n_total and toks.shape[1] == MAX_STEPS
    ms, launches = results[args.pipeline]
    value = n_total * args.steps / (ms / 1e3)
    sync_value = n_total * args.steps / (results[1][0] / 1e3)

    # ---------------- end to end with HOST buffers (`e2e`) ----------------
    # coalesce == 1: through the C ABI (gitb200_generate_host_async / _finish), one engine per batch in flight;
    # coalesce  > 1: through the Python surface (model.submit on pinned host tensors -> .cpu()), the same dynamic batching
    #                as the device-resident leg.  Either way every step's H2D pixel copy and D2H token copy is timed.
    sp = model._search_struct()
    lib = _lib.load()
    if args.coalesce > 1:
        with torch.cuda.stream(stream):
            run_device(2 * args.pipeline * args.coalesce, args.pipeline, img_host, True)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            toks = run_device(args.steps, args.pipeline, img_host, True)
            for sl in model._slots:
                if sl['stream'] is not None:
                    stream.wait_stream(sl['stream'])
            e1.record(stream)
            barrier()
            ms_e2e = e0.elapsed_time(e1)
        e2e_api = 'model.submit({image: pinned host tensor}, depth=%d, coalesce=%d) -> result().cpu()' % (args.pipeline, args.coalesce)
    else:
        n_e2e_slots = args.pipeline
        slots = []
        for k in range(n_e2e_slots):
            lib, _ = model._ensure_engine(k)
            slots.append(dict(engine=model._slots[k]['engine'],
                              stream=stream if k == 0 else torch.cuda.Stream(device=dev),
                              tok=torch.empty((B, MAX_STEPS), dtype=torch.long).pin_memory(),
                              lp=torch.empty((B,), dtype=torch.float32).pin_memory(), busy=False))
        n_out = ctypes.c_int32(0)

        def e2e_finish(sl):
            _lib.check(lib.gitb200_generate_finish(sl['engine'], ctypes.byref(n_out)), sl['engine'], 'generate_finish')
            sl['busy'] = False
            if world > 1:
                gather_captions(sl['tok'].to(dev, non_blocking=True), sl['lp'].to(dev, non_blocking=True), n_total)

        def run_host(k):
            """k steps through the C ABI with HOST buffers: H2D pixels, generate, D2H tokens; `pipeline` engines in flight."""
            for i in range(k):
                sl = slots[i % n_e2e_slots]
                if sl['busy']:
                    e2e_finish(sl)
                _lib.check(lib.gitb200_generate_host_async(sl['engine'], img_host.data_ptr(), B, 0, None, 0, ctypes.byref(sp),
                                                           sl['tok'].data_ptr(), sl['lp'].data_ptr(), sl['stream'].cuda_stream),
                           sl['engine'], 'generate_host_async')
                sl['busy'] = True
            for sl in slots:
                if sl['busy']:
                    e2e_finish(sl)

        with torch.cuda.stream(stream):
            run_host(2 * n_e2e_slots)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run_host(args.steps)
            for sl in slots[1:]:
                stream.wait_stream(sl['stream'])
            e1.record(stream)
            barrier()
            ms_e2e = e0.elapsed_time(e1)
        e2e_api = 'gitb200_generate_host_async / gitb200_generate_finish (C ABI), %d engines in flight' % n_e2e_slots
    t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_e2e = t.item()
    e2e_value = n_total * args.steps / (ms_e2e / 1e3)
    if rank == 0:
        sampler.stop()
        sampler.join(timeout=2)

    # ---------------- roofline of the dominant kernel, measured live ----------------
    peaks = measured_peaks()
    roofline = None
    extra = {}
    if rank == 0 and not args.no_micro:
        # dominant kernel = gemm_bf16_tcgen05 (profiles/: ~2/3 of the step); its largest instance is the ViT MLP
        # c_fc GEMM [B*197, 768] x [768, 3072] (+bias +QuickGELU, bf16 out): algorithmic FLOPs = 2*M*N*K.
        M, N, K = B * 197, 3072, 768
        a = (torch.randn(M, K, device=dev) * 1.0).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
        with torch.cuda.stream(stream):
            def gemm():
                rc = lib.gitb200_op_gemm(a.data_ptr(), w.data_ptr(), bias.data_ptr(), None, out.data_ptr(), M, N, K, 1, 1, 0, 1,
                                         0, stream.cuda_stream)
                assert rc == 0, _lib.last_error(None)
            for _ in range(3):
                gemm()
            durs = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                gemm()
                e1.record(stream)
                e1.synchronize()
                durs.append(e0.elapsed_time(e1))
        avg_ms = sum(durs) / len(durs)
        flops = 2.0 * M * N * K
        achieved = flops / (avg_ms / 1e3) / 1e12
        roofline = {'kernel': 'gemm_bf16_tcgen05<BN> (ViT mlp.c_fc shape %dx%dx%d, bias+QuickGELU epilogue)' % (M, N, K),
                    'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                    'frac': achieved / peaks['bf16_tflops'], 'traffic': ncu_traffic(), 'avg_launch_ms': avg_ms,
                    'peak_source': peaks['source'] + ', burst bf16 figure (kernel timed alone, L2 flushed between launches)'}
        extra['whole_step'] = {
            'algorithmic_tflop_per_step': 58.1e9 * B / 1e12,
            'achieved_tflops_whole_step': 58.1e9 * B * args.steps / (ms / 1e3) / 1e12 / world * world,
        }

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        v, sec, threads = cpu_baseline_run(12, 1, 0)
        cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': 'one batch of 12 images through oracle/git_oracle.py in as-shipped mode (no KV cache, fp32, %.1f s); '
                         'same per-image work as the batch-64 workload' % sec}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': 'GIT_BASE (CLIP ViT-B/16 + 6x768 decoder) greedy max_len=40, batch %d synthetic 224x224 '
                                   'per GPU, random-init weights' % B,
                       'global_batch': n_total, 'per_gpu_batch': B, 'parallelism': 'image-parallel x%d + 1 all_gather' % world,
                       'l2': 'inputs larger than L2: each step streams ~0.3 GB weights + 0.23 GB image K/V + activations (> 126 MB)',
                       'compute': 'bf16 operands, fp32 accumulate, fp32 residual stream',
                       'pipeline': ('%d engine launches in flight x %d submitted batches of %d per launch (model.submit: dynamic '
                                    'batching -- the batches of one launch share one encoder pass and one decode chain; the encoder / '
                                    'prefill of later launches overlap the latency-bound decode loops of earlier ones; every step does '
                                    'all of its work inside the timed region; sync_value = one batch of %d at a time through '
                                    'model(batch))' % (args.pipeline, args.coalesce, B, B))
                       if args.pipeline > 1 else '1 (synchronous model(batch) calls)'},
            'sync_value': sync_value,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'ms_per_step': ms_e2e / args.steps,
                    'h2d_bytes_per_step': img_host.numel() * 4, 'd2h_bytes_per_step': B * MAX_STEPS * 8 + B * 4, 'api': e2e_api},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
            'roofline': roofline,
            'cpu_baseline': cpu,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
