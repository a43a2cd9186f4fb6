#!/usr/bin/env python3
"""bench.py -- Mobiclip frames/sec at 400x240 (BASELINE.json `metric`), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--streams S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload ("moflex_400x240", BASELINE configs 3/5): S independent synthetic Moflex3DS streams per GPU, seeded,
I-frame every 90 frames with the keyframes of different streams staggered (so every lock-step advance carries
the steady-state mix of I- and P-pictures), 5 % intra macroblocks in P-pictures, partition trees down to 2x2,
half-pel vectors over up to five reference pictures.  A step = one lock-step advance = one new picture for each
of the S streams of a GPU.

  value  device-resident: every step's packed arrays are parsed and uploaded before the clock starts
         (mobi_batch_stage), the timed region is mobi_batch_replay only -- the reconstruction kernels.
  e2e    through the reference-facing call: HOST frame bytes in -> native entropy parse -> H2D -> reconstruct ->
         YUV->BGRA on the device -> D2H of the bitmaps into pinned host memory (mobi_batch_submit / _fetch).
  roofline   the dominant kernel (k_inter_chunk: motion compensation + dequant + inverse transforms + add/clip),
         timed alone with CUDA events on the library's own stream, against the measured HBM copy bandwidth.
  cpu_baseline / --impl reference: the reference's own decoder source compiled for the host
         (oracle/_ref, see oracle/build_ref.py), one independent stream per host thread.

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
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this command
# (profiles/): filled in after each capture, None until then.
NCU_TRAFFIC = {"k_inter_chunk": 5.68e8}   # profiles/r01l_prof_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum of one k_inter_chunk launch (414.4 + 154.1 MB)


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def gen_streams(n_streams, n_frames, first_seed, threads):
    """[stream][frame] -> bytes.  Keyframe phases are spread over the GOP."""
    from mobiclipdecoder_b200.workloads import make_stream, CONFIGS
    gop = CONFIGS[WORKLOAD][3]['gop']

    def one(i):
        s = make_stream(WORKLOAD, first_seed + i, gop_phase=(i * 37) % gop)
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


def cpu_decode_fps(streams, w, h, ver, threads, budget_s, want_bgra=True):
    """The reference decoder (oracle/_ref if built, else the oracle port) over independent streams, one per host
    thread, for about budget_s seconds of wall time (each thread replays its stream from the I-picture with a fresh
    decoder when it runs out of frames).  Returns (fps, kind, frames, seconds, threads)."""
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
        d = Dec(w, h, ver)
        gate.wait()
        stop_at = t_start[0] + budget_s
        k = done = 0
        while time.perf_counter() < stop_at:
            if k == len(fr):
                d, k = Dec(w, h, ver), 0
            if not d.decode(fr[k], 0, want_bgra)[0]:
                raise RuntimeError('CPU decoder rejected a synthetic frame')
            k += 1
            done += 1
        counts[i] = done
        return time.perf_counter()

    with cf.ThreadPoolExecutor(n) as ex:
        futs = [ex.submit(work, i) for i in range(n)]
        t_start[0] = time.perf_counter() + 0.05
        gate.wait()
        ends = [f.result() for f in futs]
    secs = max(ends) - t_start[0]
    total = sum(counts)
    return total / secs, kind, total, secs, n


def main():
    global WORKLOAD, METRIC
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--streams', type=int, default=1024, help='streams per GPU advancing in lock step')
    ap.add_argument('--threads', type=int, default=0, help='host parse threads per GPU (0 = cores / ranks)')
    ap.add_argument('--cpu-seconds', type=float, default=1.5, help='wall budget of the cpu_baseline sample (x host threads = CPU work)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
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
    cores = os.cpu_count() or 1
    threads = args.threads if args.threads > 0 else max(1, cores // max(1, world))

    from mobiclipdecoder_b200.workloads import CONFIGS
    w, h, ver, _ = CONFIGS[WORKLOAD]

    # ------------------------------------------------------------------------------------------------
    if args.impl == 'reference':
        if rank != 0:
            return 0
        from mobiclipdecoder_b200 import _build
        _build.build_mobisynth(); _build.build_oracle(); _build.build_ref()
        n_thr = cores
        per_step = 8  # frames each thread decodes per "step": a bounded sample of the workload
        n_frames = per_step * (K + Wm)
        streams = gen_streams(n_thr, n_frames, BASE_SEED, n_thr)
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import oracle_lib
        kind = 'reference' if oracle_lib.have_ref() else 'port'
        Dec = oracle_lib.Ref if kind == 'reference' else oracle_lib.Oracle
        decs = [Dec(w, h, ver) for _ in range(n_thr)]

        def step(k):
            def one(i):
                for j in range(per_step):
                    if not decs[i].decode(streams[i][k * per_step + j], 0, True)[0]:
                        raise RuntimeError('reference rejected a synthetic frame')
            list(ex.map(one, range(n_thr)))

        with cf.ThreadPoolExecutor(n_thr) as ex:
            for k in range(Wm):
                step(k)
            t0 = time.perf_counter()
            for k in range(Wm, Wm + K):
                step(k)
            secs = time.perf_counter() - t0
        fps = n_thr * per_step * K / secs
        sample = '%d host threads x %d frames of %s per step, full DecodeFrame incl. YUV->RGB' % (n_thr, per_step, WORKLOAD)
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': K, 'warmup': Wm,
            'ms_per_step': secs * 1e3 / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32 (+f32 RGB)',
            'data': 'synthetic', 'config': {'workload': WORKLOAD, 'width': w, 'height': h, 'frames_per_step': n_thr * per_step},
            'cpu_baseline': {'value': fps, 'unit': UNIT, 'cores': n_thr, 'kind': kind, 'sample': sample},
            'e2e': {'value': fps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0,
        }))
        return 0

    # ------------------------------------------------------------------------------------------------
    import numpy as np
    import torch
    import torch.distributed as dist
    from mobiclipdecoder_b200 import MobiBatch, sharding

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl native needs a CUDA device; there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    S = args.streams
    n_frames = Wm + K
    t_gen = time.time()
    streams = gen_streams(S, n_frames, BASE_SEED + rank * S, threads)
    t_gen = time.time() - t_gen
    bitstream_bytes = sum(len(f) for st in streams for f in st)

    # nvidia-smi is started well before the timed regions: its start-up (NVML initialisation over every GPU of the node)
    # takes driver locks that can hold up kernel launches of ANY rank for a millisecond or more -- seen once as a 20 %
    # slower value leg on one of eight ranks when it was started right in front of the 7 ms timed region
    sampler = ClockSampler(local_rank) if rank == 0 and not args.profile else None
    batch = MobiBatch(w, h, ver, S, device=local_rank, n_threads=threads)
    ext = torch.cuda.ExternalStream(batch.cuda_stream(), device=dev)

    # ---- value leg: staged, device-resident replay ---------------------------------------------------
    for k in range(n_frames):
        batch.stage([streams[s][k] for s in range(S)])
    batch.sync()
    staged_h2d = batch.stats()['h2d_bytes']
    batch.reset()
    if not args.profile:
        # one complete untimed pass first: a fresh box needs tens of milliseconds of work before clocks, TLBs and the
        # lazily loaded kernel images settle (measured: the first K steps after process start run ~50 % slower)
        for _ in range(2):
            batch.replay(0, n_frames)
            batch.sync()
            batch.reset()
    batch.clear_stats()
    batch.replay(0, Wm)
    batch.sync()
    if args.profile:
        batch.replay(Wm, K)
        batch.sync()
        print(json.dumps({'profile': True, 'steps': K, 'streams': S}))
        return 0
    if sampler:
        sampler.mark()
    st0 = batch.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sharding.barrier(dist if world > 1 else None, local_rank)
    torch.cuda.synchronize()
    e0.record(ext)
    batch.replay(Wm, K)
    e1.record(ext)
    batch.sync()
    torch.cuda.synchronize()
    sharding.barrier(dist if world > 1 else None, local_rank)
    ms_local = e0.elapsed_time(e1)
    if os.environ.get('MOBI_BENCH_DEBUG'):
        sys.stderr.write('rank %d: value leg %.4f ms for %d steps\n' % (rank, ms_local, K))
    st1 = batch.stats()
    ms = sharding.max_over_ranks(dist if world > 1 else None, ms_local, torch, dev)
    d = {k: st1[k] - st0[k] for k in st1}
    launches_value = d['launches']
    value = world * S * K / (ms * 1e-3)

    # ---- roofline of the dominant kernel: k_inter, per launch, CUDA events on the launching stream -------------
    # The library brackets every kernel launch with events on ITS stream (mobi_batch_set_kernel_timing) while the
    # same K steps are replayed once more.  Algorithmic bytes per launch (DESIGN.md "Kernels"): per inter MB 384 B
    # of reference picture read + 384 B written + its 16 B descriptor, 8 B per partition, 4 B per coefficient.
    peak, peak_src = load_peaks()
    batch.reset()
    batch.replay(0, Wm)
    batch.sync()
    batch.set_kernel_timing(True)
    batch.replay(Wm, K)
    kt = batch.kernel_times()
    batch.set_kernel_timing(False)
    fused = os.environ.get('MOBI_INTER_KERNEL') != 'split'   # MC + residual in one kernel (the default) or as k_mc + k_res
    inter_name = ('k_inter_v3' if os.environ.get('MOBI_INTER_KERNEL') == 'v3' else 'k_inter_chunk') if fused else 'k_mc'
    # MC = reference read once + reconstruction written + descriptor + 8 B per partition; the residual side
    # information (4 B per coefficient) belongs to whichever kernel consumes it
    mc_bytes = (768 + 16) * d['inter_mbs'] + 8 * d['parts']
    res_bytes = 16 * d['inter_mbs'] + 4 * d['inter_coefs']
    inter_bytes = mc_bytes + (4 * d['inter_coefs'] if fused else 0)
    intra_bytes = (384 + 32) * d['intra_mbs'] + 4 * d['ops'] + 4 * (d['coefs'] - d['inter_coefs'])
    n_il = max(1, kt['inter_launches'])
    inter_ms = kt['inter_ms'] / n_il
    res_ms = kt['res_ms'] / max(1, kt['res_launches'])
    achieved = (inter_bytes / n_il) / (inter_ms * 1e-3) / 1e9 if inter_ms > 0 else 0.0
    roofline = {'bound': 'hbm', 'kernel': inter_name, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'traffic': NCU_TRAFFIC.get(inter_name), 'algorithmic_bytes_per_launch': inter_bytes / n_il, 'launch_ms': inter_ms,
                'launches_timed': kt['inter_launches'], 'peak_source': peak_src,
                'step_ms_by_kernel': {inter_name: kt['inter_ms'] / K, 'k_res': kt['res_ms'] / K, 'k_intra_p_pictures': kt['intra_ms'] / K,
                                      'k_intra_i_pictures_side_stream': kt['key_ms'] / K},
                'intra_algorithmic_bytes_per_step': intra_bytes / K}
    if not fused and res_ms > 0:
        # the whole inter path (k_mc + k_res) against the fused accounting of earlier rounds, and k_res on its own side information
        roofline['inter_path'] = {'kernels': 'k_mc + k_res', 'launch_ms': inter_ms + res_ms, 'algorithmic_bytes_per_launch': (mc_bytes + 4 * d['inter_coefs']) / n_il,
                                  'achieved': ((mc_bytes + 4 * d['inter_coefs']) / n_il) / ((inter_ms + res_ms) * 1e-3) / 1e9}
        roofline['inter_path']['frac'] = roofline['inter_path']['achieved'] / peak
        roofline['k_res'] = {'launch_ms': res_ms, 'side_info_bytes_per_launch': res_bytes / n_il}

    # ---- e2e leg: host bytes -> parse -> H2D -> reconstruct -> convert -> D2H (pinned) -------------------------
    # Headline output is what the reference call returns, the BGRA bitmap (MD:260-323); the I420 variant (decoded
    # planes only, 2.7x fewer bytes over PCIe) is reported beside it.
    e2e = None
    e2e_variants = {}
    if not args.no_e2e:
        # argument tables are marshalled once, before the clock starts: the frame bytes themselves stay in ordinary host
        # memory and are read by the parser inside the timed region
        packed = [batch.pack_inputs([streams[s][k] for s in range(S)]) for k in range(n_frames)]

        def run_e2e(fmt, label):
            batch.reset_streams()
            batch.clear_staged()
            batch.clear_stats()
            for k in range(Wm):
                batch.submit(packed[k], fmt=fmt)
                batch.fetch(copy=False)
            batch.sync()
            s0 = batch.stats()
            sharding.barrier(dist if world > 1 else None, local_rank)
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
            ms_ = sharding.max_over_ranks(dist if world > 1 else None, max(ev_ms, wall_ms), torch, dev)
            return {'value': world * S * K / (ms_ * 1e-3), 'unit': UNIT,
                    'h2d_bytes_per_step': (s1['h2d_bytes'] - s0['h2d_bytes']) / K, 'd2h_bytes_per_step': (s1['d2h_bytes'] - s0['d2h_bytes']) / K,
                    'ms_per_step': ms_ / K, 'output': label, 'host_threads': threads,
                    'bitstream_bytes_per_step': bitstream_bytes / n_frames, 'gpu_launches': s1['launches'] - s0['launches']}

        e2e = run_e2e(MobiBatch.OUT_BGRA, 'BGRA bitmaps (W*H*4 per frame) in pinned host memory')
        e2e_variants['i420'] = run_e2e(MobiBatch.OUT_I420, 'tight I420 planes (W*H*3/2 per frame) in pinned host memory')
    clocks = sampler.stop() if sampler else None

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        fps, kind, total, secs, n_thr = cpu_decode_fps(streams, w, h, ver, cores, args.cpu_seconds)
        cpu = {'value': fps, 'unit': UNIT, 'cores': n_thr, 'kind': kind,
               'sample': '%d frames of %s in %.1f s: one independent stream per host thread, full DecodeFrame incl. YUV->RGB' % (total, WORKLOAD, secs)}

    if rank == 0:
        out = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': Wm, 'ms_per_step': ms / K,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/int32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'width': w, 'height': h, 'version': ver.name, 'streams_per_gpu': S, 'frames_per_step': S * world,
                       'gop': CONFIGS[WORKLOAD][3].get('gop'), 'l2_policy': 'inputs larger than L2: each step touches %.0f MB of pictures per GPU' % (2 * S * (256 if w <= 256 else 512 if w <= 512 else 1024) * h * 1.5 / 1e6),
                       'mix_per_step': {'inter_mbs': d['inter_mbs'] / K, 'intra_mbs': d['intra_mbs'] / K, 'partitions': d['parts'] / K, 'coefs': d['coefs'] / K}},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'e2e_variants': e2e_variants, 'gpu_launches': launches_value, 'clocks': clocks,
            'host': {'cores': cores, 'parse_threads_per_gpu': threads, 'stream_generation_s': t_gen, 'staged_h2d_bytes': staged_h2d},
        }
        print(json.dumps(out))
    batch.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
