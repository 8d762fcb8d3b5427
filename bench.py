#!/usr/bin/env python
"""bench.py -- the headline benchmark of the dense flow-field hot path on B200.

Workload (BASELINE.json configs[1], SURVEY.md section 8(d) cfg 2): backward warp of a 1080p batch of 64 by a
'target'-referenced flow with validity tracking, forward + backward, fp32 --

    out, valid = Flow(flow, 't', mask).apply(img, tmask, return_valid_area=True);  out.backward(g)

Metric: Mpix/s, flow-field pixels N*H*W per step.  One "step" is that call sequence on one batch.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B]

* `value`   : inputs resident in HBM, CUDA-event timed, max over ranks.
* `e2e`     : same call through the public API with HOST (pinned) buffers: every step copies flow, image, both masks
              and the upstream gradient host->device and the warped image, validity mask and both gradients
              device->host inside the timed region.
* `roofline`: the dominant kernel's algorithmic bytes / its CUDA-event duration measured live in the timed region,
              against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* `cpu_baseline`: the CPU oracle (a C/OpenMP port of the reference's CPU path, oracle/) on a bounded sample.
* `--impl reference`: only the CPU arm, as its own JSON line.

Multi-GPU: one process per GPU (torchrun), the batch dimension is sharded, every rank runs the same per-GPU
workload on its own shard (weak scaling), no collective on the data path; the time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, C = 1080, 1920, 3
BYTES_PER_PX = {  # algorithmic bytes per flow pixel, SURVEY.md section 8(d), C = 3, with both masks
    'ofb_warp_t_fwd': 8 * C + 10 + 1,      # flow 8 + payload 4C + 2 masks -> out 4C + valid 1      = 35
    'ofb_warp_t_bwd': 12 * C + 16,         # gout 4C + flow 8 + payload 4C -> gflow 8 + gpayload 4C = 52
    'ofb_flow_stats': 8 + 1,               # flow 8 + mask 1
    'ofb_flow_stats_batch': 8 + 1,
}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(p) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def python_reference_timing():
    """The UNMODIFIED Python reference timed on the build container's CPU (baseline/time_reference_cpu.py): context
    for the C/OpenMP port that `cpu_baseline` times on the GPU box (the reference itself cannot travel there)."""
    try:
        with open(os.path.join(ROOT, 'baseline', 'reference_cpu_timing.json')) as fh:
            d = json.load(fh)
        return {'what': d['reference'], 'cores': d['cores'], 'where': d['where'],
                'Mpix/s': {k: v['Mpix/s'] for k, v in d['results'].items()}}
    except Exception:
        return None


def known_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, 'profiles', 'traffic.json')
    try:
        with open(p) as fh:
            return json.load(fh).get(name)
    except Exception:
        return None


def smooth_flow(n, h, w, sigma, seed, device):
    g = torch.Generator(device='cpu').manual_seed(seed)
    lo = (torch.randn(n, 2, h // 64 + 2, w // 64 + 2, generator=g) * sigma).to(device)
    return torch.nn.functional.interpolate(lo, size=(h, w), mode='bicubic', align_corners=True).contiguous()


def make_inputs(n, device, seed0=2, sigma=8.0):
    """SURVEY.md 8(d) cfg 2: smooth-random flow sigma = 8 px, uniform image, holed masks; synthetic."""
    flow = smooth_flow(n, H, W, sigma, seed0, device)
    g = torch.Generator(device=device).manual_seed(seed0 + 1)
    img = torch.rand(n, C, H, W, generator=g, device=device)
    gout = torch.rand(n, C, H, W, generator=g, device=device)
    mask = torch.ones(n, H, W, dtype=torch.bool, device=device)
    mask[:, 300:500, 700:1100] = False
    mask[:, :, ::97] = False
    tmask = mask.roll(131, dims=2)
    return flow, img, mask, tmask, gout


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        # nvidia-smi is started BEFORE the warm-up (its first sample takes ~0.1 s, as long as the whole timed region);
        # only the samples that arrived inside the timed region count (a sample describes the 20 ms before it arrives)
        rows = self.rows
        if self.t0 is not None and self.t1 is not None:
            inside = [r for r in rows if self.t0 <= r[0] <= self.t1 + 0.03]
            rows = inside if inside else rows[-3:]
        for _t, r in rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nme, val in zip(names, f[2:6]):
                if val.lower().startswith('active'):
                    reasons.add(nme)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_arm(steps, warmup, sample_n):
    """The reference's CPU path, as restated by the oracle (C + OpenMP over all host cores), on a bounded sample of
    the same workload: batch `sample_n` instead of 64 (the metric is per pixel and flat in N)."""
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    flow, img, mask, tmask, gout = make_inputs(sample_n, 'cpu')
    times = []
    for it in range(warmup + steps):
        f = flow.clone().requires_grad_()
        i = img.clone().requires_grad_()
        t0 = time.perf_counter()
        out, valid = orc.OFlow(f, 't', mask).apply(i, tmask, return_valid_area=True)
        out.backward(gout)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    px = sample_n * H * W
    return {'value': px / (ms * 1e-3) / 1e6, 'ms_per_step': ms, 'cores': cores, 'sample_n': sample_n}


# ------------------------------------------------------------------------------------------------ GPU arm
def gpu_step(of, flow, img, mask, tmask, gout):
    f = flow.requires_grad_()
    i = img.requires_grad_()
    out, valid = of.Flow(f, 't', mask).apply(i, tmask, return_valid_area=True)
    out.backward(gout)
    gf, gi = f.grad, i.grad
    f.grad = None
    i.grad = None
    return out, valid, gf, gi


def _timeit(fn, steps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def _entry(ms, px, bytes_per_px, peak):
    return {'ms': round(ms, 3), 'Mpix/s': round(px / ms / 1e3, 1), 'bytes_per_px': bytes_per_px,
            'algorithmic_GB': round(bytes_per_px * px / 1e9, 3),
            'frac_of_hbm': round(bytes_per_px * px / (ms * 1e-3) / 1e9 / peak, 3)}


def holed(n, h, w, dev, scale=1):
    m = torch.ones(n, h, w, dtype=torch.bool, device=dev)
    m[:, 300 * scale:500 * scale, 700 * scale:1100 * scale] = False
    m[:, :, ::97] = False
    return m


def workload_combine_1080p(of, dev, peak, n=64):
    """The other half of BASELINE.json's metric: of.combine_flows forward + backward at 64 x 1080p (tensor API,
    reference flow_operations.py:84-188): mode 3 for both references (one fused W launch) and mode 2 for ref 's'
    (splat-backed: S(v1, v2 - v1), reference flow_class.py:1768).  Bytes per pixel: SURVEY.md 8(d)."""
    res = {}
    fa, fb = smooth_flow(n, H, W, 8.0, 20, dev), smooth_flow(n, H, W, 8.0, 21, dev)
    g = torch.Generator(device=dev).manual_seed(22)
    gout = torch.rand(n, 2, H, W, generator=g, device=dev)
    px = n * H * W

    def make(mode, ref):
        def fn():
            a = fa.detach().requires_grad_()
            b = fb.detach().requires_grad_()
            out = of.combine_flows(a, b, mode, ref)
            out.backward(gout)
        return fn

    # tensor API: all-true masks, so mode 3 moves 8+8 -> 8 forward and 8+8+8 -> 8+8 backward = 24 + 40
    res['combine_flows mode 3 ref t, fwd+bwd, %dx1080x1920' % n] = _entry(_timeit(make(3, 't')), px, 24 + 40, peak)
    res['combine_flows mode 3 ref s, fwd+bwd, %dx1080x1920' % n] = _entry(_timeit(make(3, 's')), px, 24 + 40, peak)
    # mode 2 ref 's': one subtraction (24) + S with a 2-channel payload, fwd 8+8 -> 8 and bwd per SURVEY 8(d) 16C+21 = 53
    res['combine_flows mode 2 ref s (splat-backed), fwd+bwd, %dx1080x1920' % n] = _entry(
        _timeit(make(2, 's')), px, 24 + 24 + 53 + 24, peak)
    return res


def workload_cfg3(of, dev, peak, n=32):
    """BASELINE.json configs[2]: forward 's' warp through the splat, 1080p batch 32, with validity; fwd and fwd+bwd."""
    res = {}
    flow = smooth_flow(n, H, W, 8.0, 5, dev)
    flow[:, :, 200:440, 300:700] = 0.0          # 5 % of the pixels exactly zero: the occlusion rule
    g = torch.Generator(device=dev).manual_seed(6)
    img = torch.rand(n, C, H, W, generator=g, device=dev)
    gout = torch.rand(n, C, H, W, generator=g, device=dev)
    mask = holed(n, H, W, dev)
    tmask = mask.roll(131, dims=2)
    px = n * H * W

    def s_fwd():
        with torch.no_grad():
            of.Flow(flow, 's', mask).apply(img, tmask, return_valid_area=True)

    def s_fwdbwd():
        f = flow.detach().requires_grad_()
        i = img.detach().requires_grad_()
        o, _v = of.Flow(f, 's', mask).apply(i, tmask, return_valid_area=True)
        o.backward(gout)

    res['cfg3 Flow.apply s + valid area, fwd, %dx3x1080x1920' % n] = _entry(_timeit(s_fwd), px, 35, peak)
    res['cfg3 Flow.apply s + valid area, fwd+bwd, %dx3x1080x1920' % n] = _entry(_timeit(s_fwdbwd), px, 35 + 69, peak)
    return res


def workload_cfg4(of, dev, peak, n=16):
    """BASELINE.json configs[3]: combine mode 3 + switch_ref on 4K batch 16 with masks, both references."""
    res = {}
    hh, ww = 2160, 3840
    fa, fb = smooth_flow(n, hh, ww, 8.0, 8, dev), smooth_flow(n, hh, ww, 8.0, 9, dev)
    m = holed(n, hh, ww, dev, 2)
    px = n * hh * ww
    for ref in 'ts':
        a, b = of.Flow(fa, ref, m), of.Flow(fb, ref, m)

        def comb():
            with torch.no_grad():
                return a.combine_with(b, 3)

        def sw():
            with torch.no_grad():
                return a.switch_ref()

        def both():
            with torch.no_grad():
                return a.combine_with(b, 3).switch_ref()

        res['cfg4 combine_with mode 3 ref %s, %dx2160x3840, masked' % (ref, n)] = _entry(_timeit(comb), px, 27, peak)
        res['cfg4 switch_ref from %s, %dx2160x3840, masked' % (ref, n)] = _entry(_timeit(sw), px, 18, peak)
        res['cfg4 combine_with mode 3 then switch_ref, ref %s, %dx2160x3840' % (ref, n)] = _entry(
            _timeit(both), px, 27 + 18, peak)
    return res


CHAIN_FRAMES = 256


def workload_cfg5(of, dev, peak, rank, world, frames_total=CHAIN_FRAMES, steps=2):
    """BASELINE.json configs[4], STRONG scaling: 256 frames of 4K flow (seeds 100..355, sigma 4 px), chained
    acc = acc.combine_with(f_i, 3); img_i = acc.apply(img, return_valid_area=True); loss = sum_i img_i.sum(); one
    backward through the chain (reference flow_class.py:1648-1810 + :755-959).  Rank g of G owns frames
    [256 g / G, 256 (g+1) / G) and chains them among themselves; no collective on the data path.
    Returns this rank's ms per chain; the caller takes the max over ranks."""
    hh, ww = 2160, 3840
    lo, hi = frames_total * rank // world, frames_total * (rank + 1) // world
    flows = [smooth_flow(1, hh, ww, 4.0, 100 + i, dev) for i in range(lo, hi)]
    g = torch.Generator(device=dev).manual_seed(7)
    img = torch.rand(1, C, hh, ww, generator=g, device=dev)

    def chain():
        fl = [f.detach().requires_grad_() for f in flows]
        im = img.detach().requires_grad_()
        acc, loss = None, None
        for f in fl:
            cur = of.Flow(f, 't')
            acc = cur if acc is None else acc.combine_with(cur, 3)
            out, _valid = acc.apply(im, return_valid_area=True)
            loss = out.sum() if loss is None else loss + out.sum()
        loss.backward()

    ms = _timeit(chain, steps=steps, warm=1)
    of.check_deferred()

    # The same chain captured ONCE into a CUDA graph (forward and backward; possible because 'deferred' validation
    # leaves no host read inside) and replayed: every kernel runs again on every replay, only the Python / autograd
    # dispatch is gone.  Reported beside the eager number, never instead of it.
    ms_graph = None
    if of.get_validation() == 'deferred':
        try:
            fl = [f.detach().requires_grad_() for f in flows]
            im = img.detach().requires_grad_()

            def body():
                acc, loss = None, None
                for f in fl:
                    cur = of.Flow(f, 't')
                    acc = cur if acc is None else acc.combine_with(cur, 3)
                    out, _valid = acc.apply(im, return_valid_area=True)
                    loss = out.sum() if loss is None else loss + out.sum()
                loss.backward()

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()
            torch.cuda.current_stream().wait_stream(side)
            for t in fl + [im]:
                t.grad = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            ms_graph = _timeit(graph.replay, steps=steps + 1, warm=1)
            of.check_deferred()
            del graph, fl, im
        except Exception as exc:        # noqa: BLE001 -- the eager number stands on its own
            ms_graph = 'capture failed: %s' % str(exc)[:120]
            torch.cuda.synchronize()
    del flows
    return ms, hi - lo, ms_graph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='per-GPU batch (the named config is 64)')
    ap.add_argument('--cpu-sample', type=int, default=4, help='batch of the bounded CPU sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--e2e-chunk', type=int, default=4, help='batch items per pipelined chunk of the e2e path')
    ap.add_argument('--validation', default='deferred', choices=['eager', 'deferred'],
                    help="flow validation mode of the package (eager = one host read per user flow, the reference's contract)")
    ap.add_argument('--no-workloads', action='store_true', help='skip combine_flows and configs 3, 4, 5')
    ap.add_argument('--chain-frames', type=int, default=CHAIN_FRAMES, help='frames of the config-5 chain (256 = the named config)')
    ap.add_argument('--sigma', type=float, default=8.0, help='flow magnitude (px) of the synthetic field; 8 = the named config')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else max(args.warmup, 1)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    workload = 'Flow.apply ref t + target mask + valid area, fwd+bwd (grads to flow and image), %dx%dx%dx%d fp32' % (
        args.batch, C, H, W)

    if args.impl == 'reference':
        if rank != 0:
            return
        r = cpu_reference_arm(args.steps, args.warmup, args.cpu_sample)
        line = {
            'impl': 'reference', 'metric': 'Mpix/s apply_flow fwd+bwd', 'value': r['value'], 'unit': 'Mpix/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'sample': 'batch %d of 64 per step, same H, W, C' % r['sample_n']},
            'cpu_baseline': {'value': r['value'], 'unit': 'Mpix/s', 'cores': r['cores'], 'kind': 'port',
                             'sample': 'oracle (C/OpenMP restatement of the reference CPU path), batch %d x 1080p, '
                                       'fwd+bwd, mean of %d steps' % (r['sample_n'], args.steps)},
            'e2e': {'value': r['value'], 'unit': 'Mpix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
        }
        print(json.dumps(line))
        return

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    import oflibpytorch_b200 as of
    from oflibpytorch_b200 import _native, _ops
    import torch.distributed as dist

    # 'deferred' validation (the opt-out of the per-flow host read, oflibpytorch_b200/utils.py): the finite test and
    # the reference's zero-flow early-outs are evaluated on the device; the deferred finite test is read back once
    # after each timed loop.  Results are identical to 'eager' (tests/test_gpu_deferred.py).
    of.set_validation(args.validation)
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.batch
    px = n * H * W
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()              # long before the timed region: nvidia-smi needs ~0.1 s for its first sample
    flow, img, mask, tmask, gout = make_inputs(n, dev, seed0=2 + 10 * rank, sigma=args.sigma)

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(args.warmup):
        gpu_step(of, flow, img, mask, tmask, gout)
    barrier()
    _native.launch_count_reset()
    _ops.PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        gpu_step(of, flow, img, mask, tmask, gout)
    e1.record()
    barrier()
    sampler.mark_end()
    of.check_deferred()
    clocks = sampler.stop() if rank == 0 else None
    launches = _native.launch_count()
    prof, _ops.PROFILE = _ops.PROFILE, None
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    per_kernel = {}
    for name, a, b in prof:
        base = name[:-3] if name.endswith('_ex') else (name[:-2] if name.endswith('_d') else name)
        per_kernel.setdefault(base, []).append(a.elapsed_time(b))
    peak, peak_src = measured_peak()
    kernels = {}
    for name, ts in per_kernel.items():
        avg = sum(ts) / len(ts)
        byts = BYTES_PER_PX.get(name, 0) * px
        kernels[name] = {'ms': avg, 'calls_per_step': len(ts) / args.steps, 'algorithmic_bytes': byts,
                         'gbs': byts / (avg * 1e-3) / 1e9}
    dominant = max(kernels, key=lambda k: kernels[k]['ms'] * kernels[k]['calls_per_step'])
    dk = kernels[dominant]
    roofline = {'bound': 'hbm', 'kernel': dominant, 'achieved': dk['gbs'], 'peak': peak, 'unit': 'GB/s',
                'frac': dk['gbs'] / peak, 'traffic': known_traffic(dominant), 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': dk['algorithmic_bytes'], 'ms_per_launch': dk['ms'],
                'share_of_step': dk['ms'] * dk['calls_per_step'] / ms,
                'all_kernels': {k: {'ms': round(v['ms'], 4), 'gbs': round(v['gbs'], 1),
                                    'frac': round(v['gbs'] / peak, 4)} for k, v in kernels.items()},
                'step_frac': (sum(BYTES_PER_PX[k] for k in ('ofb_warp_t_fwd', 'ofb_warp_t_bwd')) * px /
                              (ms * 1e-3) / 1e9) / peak}

    # ---------------------------------------------------------------- end to end with host buffers
    e2e = None
    if not args.no_e2e:
        # pinned host copies of the whole batch (9 GB per rank at the named config); if any rank cannot pin that much
        # (8 ranks on one host), every rank falls back to a quarter of the batch -- agreed collectively so that the
        # barriers below stay matched
        from oflibpytorch_b200 import numa
        numa_info = numa.bind_to_gpu_node(local_rank)     # pin from the GPU's own NUMA node (no-op on one node)
        e2e_n, host_in, host_out = n, None, None
        for attempt_n in (n, max(1, n // 4)):
            ok = 1
            try:
                host_in = [x[:attempt_n].detach().cpu().pin_memory() for x in (flow, img, mask, tmask, gout)]
                o, v, gf, gi = gpu_step(of, flow[:attempt_n].detach(), img[:attempt_n].detach(), mask[:attempt_n],
                                        tmask[:attempt_n], gout[:attempt_n])
                host_out = [torch.empty(x.shape, dtype=x.dtype).pin_memory() for x in (o, v, gf, gi)]
                del o, v, gf, gi
            except (RuntimeError, MemoryError):
                ok, host_in, host_out = 0, None, None
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                e2e_n = attempt_n
                break
            host_in = host_out = None
        if host_in is None:
            raise RuntimeError('bench.py: could not pin the host buffers of the end-to-end path on every rank')
        px_e2e = e2e_n * H * W
        h2d = sum(x.numel() * x.element_size() for x in host_in)
        d2h = sum(x.numel() * x.element_size() for x in host_out)

        from oflibpytorch_b200.streaming import HostWarpPipeline
        pipe = HostWarpPipeline(dev, chunk=args.e2e_chunk)

        def e2e_step():
            # public host-buffer entry point: chunked H2D -> Flow.apply + backward -> D2H, three streams overlapped
            pipe.run(host_in[0], host_in[1], 't', host_in[2], host_in[3], host_in[4],
                     host_out[0], host_out[1], host_out[2], host_out[3])

        for _ in range(2):
            e2e_step()
        barrier()
        k2 = max(3, min(args.steps, 5))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(k2):
            e2e_step()
        s1.record()
        barrier()
        t2 = torch.tensor([s0.elapsed_time(s1) / k2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e = {'value': world * px_e2e / (float(t2.item()) * 1e-3) / 1e6, 'unit': 'Mpix/s',
               'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': float(t2.item()),
               'steps': k2, 'batch_per_gpu': e2e_n,
               'how': 'oflibpytorch_b200.streaming.HostWarpPipeline, chunk %d, pinned host buffers, '
                      'H2D/compute/D2H on three streams' % args.e2e_chunk,
               'numa': numa_info}
        del host_in, host_out

    # ---------------------------------------------------------------- the rest of BASELINE.json's metric and configs
    # combine_flows fwd+bwd at 64 x 1080p, configs 3 and 4 (rank 0, N = 1), and config 5 as STRONG scaling over all
    # ranks: 256 frames of 4K sharded [256 g / G, 256 (g+1) / G), time = max over ranks.
    workloads = None
    if not args.no_workloads:
        del flow, img, mask, tmask, gout
        torch.cuda.empty_cache()
        workloads = {}
        if world == 1:
            workloads.update(workload_combine_1080p(of, dev, peak))
            torch.cuda.empty_cache()
            workloads.update(workload_cfg3(of, dev, peak))
            torch.cuda.empty_cache()
            workloads.update(workload_cfg4(of, dev, peak))
            torch.cuda.empty_cache()
        barrier()
        ms5, mine, ms5g = workload_cfg5(of, dev, peak, rank, world, args.chain_frames)
        t5 = torch.tensor([ms5, ms5g if isinstance(ms5g, float) else -1.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        ms5 = float(t5[0].item())
        px5 = args.chain_frames * 2160 * 3840
        e5 = _entry(ms5, px5, 155, peak * world)
        if isinstance(ms5g, float) and float(t5[1].item()) > 0:
            g5 = _entry(float(t5[1].item()), px5, 155, peak * world)
            e5['cuda_graph_replay'] = {'ms': g5['ms'], 'Mpix/s': g5['Mpix/s'], 'frac_of_hbm': g5['frac_of_hbm'],
                                       'what': 'the same chain (fwd + bwd) captured once with torch.cuda.graph and '
                                               'replayed; all kernels run per replay, Python dispatch removed'}
        elif ms5g is not None:
            e5['cuda_graph_replay'] = ms5g
        e5.update({'frames_total': args.chain_frames, 'frames_per_rank': mine, 'n_gpus': world, 'scaling': 'strong',
                   'what': 'acc = acc.combine_with(f_i, 3); acc.apply(img, return_valid_area=True); one backward '
                           'through the chain; time = max over ranks; frac is of n_gpus x the HBM peak'})
        workloads['cfg5 chained combine+apply fwd+bwd, %d frames of 2160x3840, strong scaling' % args.chain_frames] = e5

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_arm(3, 1, args.cpu_sample)
        cpu = {'value': r['value'], 'unit': 'Mpix/s', 'cores': r['cores'], 'kind': 'port',
               'sample': 'oracle (C/OpenMP restatement of the reference CPU path), batch %d x 1080p, fwd+bwd, '
                         'mean of 3 steps; the metric is per flow pixel and flat in the batch size, so the sample '
                         'runs batch %d where the GPU arm runs %d' % (r['sample_n'], r['sample_n'], n),
               'python_reference': python_reference_timing()}

    if rank == 0:
        line = {
            'metric': 'Mpix/s apply_flow fwd+bwd', 'value': world * px / (ms_max * 1e-3) / 1e6, 'unit': 'Mpix/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_max,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'per_gpu_batch': n, 'global_batch': n * world,
                       'validation': args.validation,
                       'parallelism': 'batch shards, one per GPU, no collective',
                       'l2': 'inputs (%.1f GB per step) exceed the 126 MB L2; no flush needed' % (
                           (BYTES_PER_PX['ofb_warp_t_fwd'] + BYTES_PER_PX['ofb_warp_t_bwd']) * px / 1e9)},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': launches, 'clocks': clocks,
            'workloads': workloads,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
