#!/usr/bin/env python3
"""bench.py -- Mobiclip frames/sec at 400x240 (BASELINE.json `metric`), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--streams S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload ("moflex_400x240", BASELINE configs 3/5): S independent synthetic Moflex3DS streams per GPU, seeded,
I-frame every 90 frames with the keyframes of different streams staggered (so every lock-step advance carries
the steady-state mix of I- and P-pictures), 5 % intra macroblocks in P-pictures, partition trees down to 2x2,
half-pel vectors over up to five reference pictures.  A step = one lock-step advance = one new picture for each
of the S streams of a GPU.

  value  device-resident, the WHOLE north_star path: every step's packed arrays are parsed and uploaded before the clock
         starts (mobi_batch_stage); the timed region is mobi_batch_replay_convert -- reconstruction kernels AND YUV->BGRA
         (k_bgra) -- of the K steps, repeated R times inside one pair of CUDA events (>= 0.3 s of device time).
         `extra.value_reconstruct_only` is the same without the conversion (what round 1 reported as value).
  e2e    through the reference-facing call: HOST frame bytes in -> native entropy parse -> H2D -> reconstruct ->
         YUV->BGRA on the device -> D2H of the bitmaps into pinned host memory (mobi_batch_submit / _fetch), with the
         host-side phase breakdown of the library (mobi_batch_get_phase_times).
  roofline   the dominant kernel (the inter kernel: motion compensation + dequant + inverse transforms + add/clip),
         timed alone with CUDA events on the library's own stream, against the measured HBM copy bandwidth;
         `roofline_bgra` the same for k_bgra.
  cpu_baseline / --impl reference: the reference's own decoder source compiled for the host
         (oracle/_ref, see oracle/build_ref.py), one independent stream per host thread, free-running.
  extra  BASELINE.json configs 2 and 3 as stated (1024 pre-parsed 256x192 P-frames, kernels only; ONE 400x240 stream
         end to end with per-frame latency) and the un-staggered step (every stream on an I-picture).

Multi-GPU: streams are independent, rank r owns its own S streams, there is no data-path
collective; torch.distributed is used for the barrier and for the max-over-ranks time only.
"""
import argparse
import concurrent.futures as cf
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'mobiclip_frames_per_sec_400x240'
UNIT = 'frames/s'
WORKLOAD = 'moflex_400x240'
BASE_SEED = 1000


def ncu_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on `workload`, from the committed `ncu --set full`
    capture of this command (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep): None where no
    capture of that workload is committed."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        return t.get(workload, {}).get(kernel)
    except Exception:
        return None


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def gen_streams(workload, n_streams, n_frames, first_seed, threads, stagger=True):
    """[stream][frame] -> bytes.  Keyframe phases are spread over the GOP."""
    from mobiclipdecoder_b200.workloads import make_stream, CONFIGS
    gop = CONFIGS[workload][3].get('gop', 0)

    def one(i):
        s = make_stream(workload, first_seed + i, gop_phase=((i * 37) % gop if (stagger and gop) else 0))
        out = [s.next_frame()[0] for _ in range(n_frames)]
        s.close()
        return out

    with cf.ThreadPoolExecutor(max(1, threads)) as ex:
        return list(ex.map(one, range(n_streams)))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe).  The process is started
    early (its start-up must not fall into a timed region); only the samples between the first and the last mark() --
    the start of the value leg and the end of the end-to-end legs -- are reported."""
    Q = 'timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(gpu_index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        self.marks = []

    def mark(self):
        self.marks.append(time.time())

    @staticmethod
    def _when(stamp):
        import datetime
        try:
            return datetime.datetime.strptime(stamp, '%Y/%m/%d %H:%M:%S.%f').timestamp()
        except ValueError:
            return None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.mark()
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = []
        for line in open(self.f.name):
            parts = [x.strip() for x in line.split(',')]
            if len(parts) >= 8:
                try:
                    rows.append((self._when(parts[0]), float(parts[1]), float(parts[2]), parts[4:8]))
                except ValueError:
                    pass
        os.unlink(self.f.name)
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        lo, hi = self.marks[0], self.marks[-1]
        busy = [r for r in rows if r[0] is not None and lo - 0.05 <= r[0] <= hi + 0.05]
        rows = busy if busy else rows[len(rows) // 2:]   # (timestamps unreadable: the later half covers the timed legs, stop() follows them)
        sm = sorted(r[1] for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in rows for i, v in enumerate(r[3]) if v.lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': rows[0][2], 'reasons': reasons, 'samples': len(rows)}


def cpu_decode_fps(streams, w, h, ver, threads, budget_s=None, frames_each=None, warm_each=0, want_bgra=True):
    """The reference decoder (oracle/_ref if built, else the oracle port) over independent streams, one per host thread,
    FREE-RUNNING: no barrier between frames or steps.  Either for about budget_s seconds of wall time (each thread replays
    its stream from the I-picture with a fresh decoder when it runs out of frames) or for exactly frames_each frames per
    thread after warm_each untimed ones.  Returns (fps, kind, frames, seconds, threads)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle_lib
    kind = 'reference' if oracle_lib.have_ref() else 'port'
    Dec = oracle_lib.Ref if kind == 'reference' else oracle_lib.Oracle
    n = min(threads, len(streams))
    counts = [0] * n
    gate = threading.Barrier(n + 1)
    t_start = [0.0]

    def work(i):
        fr = streams[i]
        state = {'d': Dec(w, h, ver), 'k': 0}

        def one():
            if state['k'] == len(fr):
                state['d'], state['k'] = Dec(w, h, ver), 0
            if not state['d'].decode(fr[state['k']], 0, want_bgra)[0]:
                raise RuntimeError('CPU decoder rejected a synthetic frame')
            state['k'] += 1

        for _ in range(warm_each):
            one()
        gate.wait()
        done = 0
        if frames_each is not None:
            for _ in range(frames_each):
                one()
            done = frames_each
        else:
            stop_at = t_start[0] + budget_s
            while time.perf_counter() < stop_at:
                one()
                done += 1
        counts[i] = done
        return time.perf_counter()

    with cf.ThreadPoolExecutor(n) as ex:
        futs = [ex.submit(work, i) for i in range(n)]
        while gate.n_waiting < n:   # every thread has finished its warm-up
            if any(f.done() for f in futs):
                [f.result() for f in futs if f.done()]   # (a thread failed: raise its error instead of waiting for ever)
            time.sleep(0.001)
        t_start[0] = time.perf_counter()
        gate.wait()
        ends = [f.result() for f in futs]
    secs = max(ends) - t_start[0]
    total = sum(counts)
    return total / secs, kind, total, secs, n


def workload_config(workload, S, world):
    """The `config` object: what the workload IS.  Both arms print exactly this (what a run MEASURED of it goes elsewhere)."""
    from mobiclipdecoder_b200.workloads import CONFIGS
    w, h, ver, ov = CONFIGS[workload]
    stride = 256 if w <= 256 else 512 if w <= 512 else 1024
    return {'workload': workload, 'width': w, 'height': h, 'version': ver.name, 'streams_per_gpu': S, 'frames_per_step': S * world,
            'gop': ov.get('gop'), 'l2_policy': 'inputs larger than L2: each step touches %.0f MB of pictures per GPU' % (2 * S * stride * h * 1.5 / 1e6)}


def main():
    global WORKLOAD, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--streams', type=int, default=1024, help='streams per GPU advancing in lock step')
    ap.add_argument('--threads', type=int, default=0, help='host parse threads per GPU (0 = cores / ranks)')
    ap.add_argument('--repeats', type=int, default=0, help='replays of the K staged steps inside the timed region of the value leg (0 = enough for 0.3 s)')
    ap.add_argument('--cpu-seconds', type=float, default=1.5, help='wall budget of the cpu_baseline sample (x host threads = CPU work)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-affinity', action='store_true', help='N > 1: do not pin each rank to its own slice of the host cores')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the config 2 / config 3 / un-staggered legs')
    ap.add_argument('--profile', action='store_true', help='short run for ncu: value leg only')
    ap.add_argument('--workload', default=WORKLOAD, help='synthetic workload (mobiclipdecoder_b200/workloads.py); the default is the one the metric is quoted on, '
                    'mods_256x192 (config 1) and moc5_640x480 (config 4) are reported beside it in BASELINE.md')
    args = ap.parse_args()
    if args.workload != WORKLOAD:
        from mobiclipdecoder_b200.workloads import CONFIGS as _C
        WORKLOAD = args.workload
        METRIC = 'mobiclip_frames_per_sec_%dx%d' % (_C[WORKLOAD][0], _C[WORKLOAD][1])

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    K, Wm = args.steps, max(args.warmup, 3 if args.impl == 'native' and not args.profile else args.warmup)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    threads = args.threads if args.threads > 0 else max(1, cores // max(1, world))

    from mobiclipdecoder_b200.workloads import CONFIGS
    w, h, ver, _ = CONFIGS[WORKLOAD]
    S = args.streams
    config = workload_config(WORKLOAD, S, world)

    # ------------------------------------------------------------------------------------------------
    if args.impl == 'reference':
        if rank != 0:
            return 0
        from mobiclipdecoder_b200 import _build
        _build.build_mobisynth(); _build.build_oracle(); _build.build_ref()
        # A bounded SAMPLE of the workload: one stream per host thread (the workload's first `cores` streams); a "step" is
        # STEP_S seconds of all threads decoding free-running (no barrier between frames or steps: the workload's streams are
        # independent), so the figure is total frames / wall time exactly as in the native arm's cpu_baseline sample -- a fixed
        # frame count per thread would time the slowest thread of a noisy VM instead.
        STEP_S, n_thr, per_step = 0.15, cores, 8
        streams = gen_streams(WORKLOAD, n_thr, per_step * (K + Wm) + 40, BASE_SEED, n_thr)
        fps, kind, total, secs, n_thr = cpu_decode_fps(streams, w, h, ver, n_thr, budget_s=STEP_S * K, warm_each=per_step * Wm)
        sample = '%d of the workload\'s streams, one per host thread, free-running for %d x %.2f s (%d frames); full DecodeFrame incl. YUV->RGB' % (n_thr, K, STEP_S, total)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': K, 'warmup': Wm,
            'ms_per_step': secs * 1e3 / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 (+f32 RGB)',
            'data': 'synthetic', 'config': config,
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': n_thr, 'kind': kind, 'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
        }))
        return 0

    # ------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    from mobiclipdecoder_b200 import MobiBatch, MobiclipDecoder, sharding

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        if not args.no_affinity:
            # each rank parses on its own slice of the host cores (and first-touches its pinned arenas from there)
            my_cores = sharding.pin_to_cores(local_rank, world)
            threads = args.threads if args.threads > 0 else max(1, len(my_cores))
    dev = torch.device('cuda', local_rank)
    D = dist if world > 1 else None

    n_frames = Wm + K
    t_gen = time.time()
    streams = gen_streams(WORKLOAD, S, n_frames, sharding.stream_seed(BASE_SEED, rank * S), threads)   # rank r owns global streams r*S .. r*S + S-1
    t_gen = time.time() - t_gen
    bitstream_bytes = sum(len(f) for st in streams for f in st)

    # nvidia-smi is started well before the timed regions: its start-up (NVML initialisation over every GPU of the node)
    # takes driver locks that can hold up kernel launches of ANY rank for a millisecond or more
    sampler = ClockSampler(local_rank) if rank == 0 and not args.profile else None
    batch = MobiBatch(w, h, ver, S, device=local_rank, n_threads=threads)
    ext = torch.cuda.ExternalStream(batch.cuda_stream(), device=dev)
    BGRA = MobiBatch.OUT_BGRA

    def timed(fn, b=None, stream=None):
        """Device time of fn() on the batch's stream, bracketed by barrier + synchronize on both sides; max over ranks."""
        b, stream = b or batch, stream or ext
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sharding.barrier(D, local_rank)
        torch.cuda.synchronize()
        e0.record(stream)
        fn()
        e1.record(stream)
        b.sync()
        torch.cuda.synchronize()
        sharding.barrier(D, local_rank)
        return sharding.max_over_ranks(D, e0.elapsed_time(e1), torch, dev)

    # ---- stage: every step parsed and uploaded before any clock starts ----------------------------------
    for k in range(n_frames):
        batch.stage([streams[s][k] for s in range(S)])
    batch.sync()
    staged_h2d = batch.stats()['h2d_bytes']
    batch.reset()
    if args.profile:
        batch.replay(0, Wm, BGRA)
        batch.sync()
        batch.replay(Wm, K, BGRA)
        batch.sync()
        print(json.dumps({'profile': True, 'steps': K, 'streams': S}))
        return 0
    # one complete untimed pass first: a fresh box needs tens of milliseconds of work before clocks, TLBs and the
    # lazily loaded kernel images settle (measured: the first K steps after process start run ~50 % slower)
    for _ in range(2):
        batch.replay(0, n_frames, BGRA)
        batch.sync()
        batch.reset()

    # ---- value leg: the K staged steps, R times over, reconstruction + YUV->BGRA, device-resident -----------
    # Steps Wm .. Wm+K-1 are replayed back to back R times.  A staged step holds absolute picture addresses, so every replay
    # does the same work on the same addresses (the pixel VALUES it finds in the ring differ from the second replay on --
    # nothing in these kernels is data-dependent).
    batch.replay(0, Wm, BGRA)
    batch.sync()
    ms_probe = timed(lambda: batch.replay(Wm, K, BGRA))
    R = args.repeats if args.repeats > 0 else max(50, int(300.0 / max(ms_probe, 1e-3)) + 1)
    if sampler:
        sampler.mark()
    batch.clear_stats()
    st0 = batch.stats()

    def value_leg():
        for _ in range(R):
            batch.replay(Wm, K, BGRA)
    ms_total = timed(value_leg)
    st1 = batch.stats()
    d = {k: (st1[k] - st0[k]) / R for k in st1}   # per replay of the K steps
    launches_value = st1['launches'] - st0['launches']
    ms = ms_total / R   # per K steps
    value = world * S * K / (ms * 1e-3)
    reps = sorted(timed(lambda: batch.replay(Wm, K, BGRA)) / K for _ in range(7))   # spread: a few single replays, timed one by one
    ms_rec = timed(lambda: [batch.replay(Wm, K) for _ in range(R)]) / R             # the same without the conversion

    # ---- roofline of the dominant kernels, per launch, CUDA events on the launching stream -------------------
    # The library brackets every kernel launch with events on ITS stream (mobi_batch_set_kernel_timing) while the
    # same K steps are replayed once more.  Algorithmic bytes per launch (DESIGN.md "Kernels"): per inter MB
    # 384 B of reference picture read + 384 B written + its 16 B descriptor, 8 B per partition, 4 B per coefficient;
    # YUV->BGRA: 1.5 W H read + 4 W H written per picture.
    peak, peak_src = load_peaks()
    batch.reset()
    batch.replay(0, Wm, BGRA)
    batch.sync()
    batch.set_kernel_timing(True)
    batch.replay(Wm, K, BGRA)
    kt = batch.kernel_times()
    batch.set_kernel_timing(False)
    fused = os.environ.get('MOBI_INTER_KERNEL') != 'split'   # MC + residual in one kernel (the default) or as k_mc + k_res
    inter_name = ('k_inter_v3' if os.environ.get('MOBI_INTER_KERNEL') == 'v3' else 'k_inter_chunk') if fused else 'k_mc'
    mc_bytes = (768 + 16) * d['inter_mbs'] + 8 * d['parts']
    res_bytes = 16 * d['inter_mbs'] + 4 * d['inter_coefs']
    inter_bytes = mc_bytes + (4 * d['inter_coefs'] if fused else 0)
    intra_bytes = (384 + 32) * d['intra_mbs'] + 4 * d['ops'] + 4 * (d['coefs'] - d['inter_coefs'])
    n_il = max(1, kt['inter_launches'])
    inter_ms = kt['inter_ms'] / n_il
    res_ms = kt['res_ms'] / max(1, kt['res_launches'])
    achieved = (inter_bytes / n_il) / (inter_ms * 1e-3) / 1e9 if inter_ms > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': inter_name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': ncu_traffic(WORKLOAD, inter_name), 'algorithmic_bytes_per_launch': inter_bytes / n_il, 'launch_ms': inter_ms,
                'launches_timed': kt['inter_launches'], 'peak_source': peak_src,
                'step_ms_by_kernel': {inter_name: kt['inter_ms'] / K, 'k_intra_p_pictures': kt['intra_ms'] / K,
                                      'k_intra_i_pictures_side_stream': kt['key_ms'] / K, 'k_bgra': kt['bgra_ms'] / K},
                'intra_algorithmic_bytes_per_step': intra_bytes / K}
    if not fused and res_ms > 0:
        roofline['step_ms_by_kernel']['k_res'] = kt['res_ms'] / K
        roofline['inter_path'] = {'kernels': 'k_mc + k_res', 'launch_ms': inter_ms + res_ms, 'algorithmic_bytes_per_launch': (mc_bytes + 4 * d['inter_coefs']) / n_il,
                                  'achieved': ((mc_bytes + 4 * d['inter_coefs']) / n_il) / ((inter_ms + res_ms) * 1e-3) / 1e9}
        roofline['inter_path']['frac'] = roofline['inter_path']['achieved'] / peak
        roofline['k_res'] = {'launch_ms': res_ms, 'side_info_bytes_per_launch': res_bytes / n_il}
    bgra_ms = kt['bgra_ms'] / max(1, kt['bgra_launches'])
    bgra_bytes = S * (1.5 * w * h + 4 * w * h)
    roofline_bgra = {'bound': 'hbm', 'kernel': 'k_bgra', 'achieved': bgra_bytes / (bgra_ms * 1e-3) / 1e9 if bgra_ms > 0 else 0.0, 'peak': peak, 'unit': 'GB/s',
                     'traffic': ncu_traffic(WORKLOAD, 'k_bgra'), 'algorithmic_bytes_per_launch': bgra_bytes, 'launch_ms': bgra_ms, 'launches_timed': kt['bgra_launches'],
                     'note': 'runs after the step\'s reconstruction kernels on the same stream'}
    roofline_bgra['frac'] = roofline_bgra['achieved'] / peak

    # ---- e2e leg: host bytes -> parse -> H2D -> reconstruct -> convert -> D2H (pinned) -------------------------
    # Headline output is what the reference call returns, the BGRA bitmap (MD:260-323); the I420 variant (decoded
    # planes only, 2.7x fewer bytes over PCIe) is reported beside it.
    e2e = None
    e2e_variants = {}
    if not args.no_e2e:
        # argument tables are marshalled once, before the clock starts: the frame bytes themselves stay in ordinary host
        # memory and are read by the parser inside the timed region
        packed = [batch.pack_inputs([streams[s][k] for s in range(S)]) for k in range(n_frames)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def run_e2e(fmt, label):
            batch.reset_streams()
            batch.clear_staged()
            for k in range(Wm):
                batch.submit(packed[k], fmt=fmt)
                batch.fetch(copy=False)
            batch.sync()
            batch.clear_stats()
            s0 = batch.stats()
            sharding.barrier(D, local_rank)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record(ext)
            batch.submit(packed[Wm], fmt=fmt)
            for k in range(Wm + 1, Wm + K):
                batch.submit(packed[k], fmt=fmt)   # host parses step k while the GPU still works on step k-1
                batch.fetch(copy=False)
            batch.fetch(copy=False)
            e1.record(ext)
            batch.sync()
            wall_ms = (time.perf_counter() - t0) * 1e3
            ev_ms = e0.elapsed_time(e1)
            s1 = batch.stats()
            ph = batch.phase_times()
            ms_ = sharding.max_over_ranks(D, max(ev_ms, wall_ms), torch, dev)
            per_rank = sharding.gather_objects(D, dict({k_: round(v / K, 3) for k_, v in ph.items()}, wall=round(wall_ms / K, 3), threads=threads))
            return {'value': world * S * K / (ms_ * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': (s1['h2d_bytes'] - s0['h2d_bytes']) / K, 'd2h_bytes_per_step': (s1['d2h_bytes'] - s0['d2h_bytes']) / K,
                    'ms_per_step': ms_ / K, 'output': label, 'host_threads': threads,
                    'bitstream_bytes_per_step': bitstream_bytes / n_frames, 'gpu_launches': s1['launches'] - s0['launches'],
                    # where rank 0's calling thread spent the step (host wall time per step): the parse fans out over host_threads;
                    # fetch_wait is the time the host had nothing left to do but wait for the copy-back
                    'host_ms_per_step': {k_: v / K for k_, v in ph.items()},
                    'host_ms_per_step_by_rank': per_rank if world > 1 else None,
                    'd2h_gbps_per_gpu': (s1['d2h_bytes'] - s0['d2h_bytes']) / K / (ms_ / K * 1e-3) / 1e9}

        e2e = run_e2e(BGRA, 'BGRA bitmaps (W*H*4 per frame) in pinned host memory')
        e2e_variants['i420'] = run_e2e(MobiBatch.OUT_I420, 'tight I420 planes (W*H*3/2 per frame) in pinned host memory')
    if sampler:
        sampler.mark()

    # ---- BASELINE.json configs 2 and 3 as stated, and the un-staggered step (rank 0, one GPU) -----------------
    extra = {'value_reconstruct_only': {'value': world * S * K / (ms_rec * 1e-3), 'ms_per_step': ms_rec / K, 'unit': UNIT,
                                        'note': 'the value leg without k_bgra (what round 1 reported as value)'},
             'value_spread_ms_per_step': {'min': reps[0], 'median': reps[len(reps) // 2], 'max': reps[-1], 'single_replays': len(reps)}}
    if rank == 0 and world == 1 and not args.no_extras:
        # un-staggered: step 0 of every stream is its I-picture -- S pictures through the intra kernels at once
        batch.reset_streams()
        batch.clear_staged()
        for k in range(min(3, n_frames)):
            batch.stage([streams[s][k] for s in range(S)])
        batch.sync()
        batch.reset()
        batch.replay(0, 1)
        batch.sync()
        ms_key = min(timed(lambda: batch.replay(0, 1)) for _ in range(5))
        extra['unstaggered_key_step'] = {'ms': ms_key, 'pictures': S, 'vs_steady_state_step': ms_key / (ms_rec / K),
                                         'note': 'every stream on an I-picture in the same step (reconstruction only)'}
        batch.close()
        batch = None
        # config 2: 1024 pre-parsed 256x192 P-frames, IDCT + MC kernels only
        w2, h2, v2, _ = CONFIGS['pframes_256x192']
        st2 = gen_streams('pframes_256x192', 1024, 2, 1, threads, stagger=False)
        b2 = MobiBatch(w2, h2, v2, 1024, device=local_rank, n_threads=threads)
        ext2 = torch.cuda.ExternalStream(b2.cuda_stream(), device=dev)
        for k in range(2):
            b2.stage([st2[s][k] for s in range(1024)])
        b2.sync(); b2.reset()
        b2.replay(0, 2); b2.sync()
        b2.clear_stats()
        ms2 = timed(lambda: [b2.replay(1, 1) for _ in range(200)], b2, ext2) / 200
        s2 = b2.stats()
        by2 = ((768 + 16) * s2['inter_mbs'] + 8 * s2['parts'] + 4 * s2['inter_coefs']) / 200
        extra['config2_pframes_256x192'] = {'frames_per_s': 1024 / (ms2 * 1e-3), 'ms_per_1024_frames': ms2, 'inter_mbs': s2['inter_mbs'] / 200,
                                            'algorithmic_GBps': by2 / (ms2 * 1e-3) / 1e9, 'frac_of_peak': by2 / (ms2 * 1e-3) / 1e9 / peak,
                                            'note': '1024 independent (reference picture, P-frame) pairs, inter macroblocks only, packed arrays resident; 200 replays'}
        b2.close()
        # config 3: ONE 400x240 stream end to end through the reference-facing per-frame call (parse -> H2D -> kernels -> BGRA -> D2H)
        w3, h3, v3, _ = CONFIGS['moflex_400x240']
        one = gen_streams('moflex_400x240', 1, 400, 1, 1, stagger=False)[0]
        dec = MobiclipDecoder(w3, h3, v3, device=local_rank)
        lat = []
        for i, fr in enumerate(one):
            t0 = time.perf_counter()
            dec.Data, dec.Offset = fr, 0
            if dec.DecodeFrame() is None:
                raise RuntimeError('single-stream decode failed')
            if i >= 20:
                lat.append((time.perf_counter() - t0) * 1e3)
        dec.close()
        lat.sort()
        extra['config3_single_stream_400x240'] = {'frames_per_s': 1e3 * len(lat) / sum(lat), 'latency_ms': {'p50': lat[len(lat) // 2], 'p99': lat[int(len(lat) * 0.99)], 'min': lat[0]},
                                                  'frames': len(lat), 'note': 'mobi_decode_frame + mobi_read_bgra per frame, host buffers, one stream, no batching'}
        # the same stream with the bitmap of frame k - 1 fetched while frame k is parsed and reconstructed (a player that shows frame k - 1)
        b1 = MobiBatch(w3, h3, v3, 1, device=local_rank, n_threads=1)
        for fr in one[:20]:
            b1.submit([fr], fmt=BGRA); b1.fetch(copy=False)
        t0 = time.perf_counter()
        b1.submit([one[20]], fmt=BGRA)
        for fr in one[21:]:
            b1.submit([fr], fmt=BGRA)
            b1.fetch(copy=False)
        b1.fetch(copy=False)
        dt = time.perf_counter() - t0
        b1.close()
        extra['config3_single_stream_400x240']['pipelined_frames_per_s'] = (len(one) - 20) / dt
    clocks = sampler.stop() if sampler else None

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        fps, kind, total, secs, n_thr = cpu_decode_fps(streams, w, h, ver, cores, budget_s=args.cpu_seconds)
        cpu = {'value': fps, 'unit': UNIT, 'cores': n_thr, 'kind': kind,
               'sample': '%d frames of %s in %.1f s: one independent stream per host thread, free-running, full DecodeFrame incl. YUV->RGB' % (total, WORKLOAD, secs)}
        extra['cpu_single_core_frames_per_s'] = cpu_decode_fps(streams[:1], w, h, ver, 1, budget_s=min(1.0, args.cpu_seconds))[0]

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 (+f32 RGB)', 'data': 'synthetic',
            'config': config,
            'value_region': {'replays_of_the_K_steps': R, 'device_ms': ms_total, 'what': 'reconstruction + YUV->BGRA, device-resident (mobi_batch_replay_convert)'},
            'workload_mix_per_step': {'inter_mbs': d['inter_mbs'] / K, 'intra_mbs': d['intra_mbs'] / K, 'partitions': d['parts'] / K, 'coefs': d['coefs'] / K},
            'roofline': roofline, 'roofline_bgra': roofline_bgra, 'cpu_baseline': cpu, 'e2e': e2e, 'e2e_variants': e2e_variants,
            'gpu_launches': launches_value, 'clocks': clocks, 'extra': extra,
            'host': {'cores': cores, 'parse_threads_per_gpu': threads, 'stream_generation_s': t_gen, 'staged_h2d_bytes': staged_h2d},
        }
        print(json.dumps(out))
    if batch is not None:
        batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
