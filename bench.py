#!/usr/bin/env python
"""Benchmark of the pose-guided rendering hot path (rasterise -> warp -> generator -> composite).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric "rendered frames/s
(gen+warp+blend)" is defined on): one 2x-interpolation clip per step — 65 frames at 512x512,
33 key frames, 32 generated frames — with synthetic joints / key frames / flows and
seeded random-init weights with converged spectral norm.  A step renders the whole clip: 32 label
rasterisations (key-frame labels are dead inputs of the reference generator), 32 background warps, one
batch-32 generator forward, the mask blend and the uint8 conversion of all 65 frames.  Under torchrun each rank renders its own clip per step (weak scaling) and the
uint8 frames are gathered with NCCL inside the timed region.

One JSON line is printed by rank 0 (contract in the task statement).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'render-in-between_b200')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 512
N_KEY = 33
RATE = 2
T = (N_KEY - 1) * RATE + 1
GEN_FRAMES = T - N_KEY                      # 32 generated frames per clip
FLOP_PER_PIXEL = 883.5e3                    # 2*MAC of the 96 convs, SURVEY.md §8d
CONV_FLOP_PER_FRAME = FLOP_PER_PIXEL * H * W


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return p.get('bf16_tflops_sustained', 1369.2), p.get('hbm_gbs', 6548.8), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '20'], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    @staticmethod
    def _epoch(ts):
        import datetime
        try:
            return datetime.datetime.strptime(ts.strip(), '%Y/%m/%d %H:%M:%S.%f').timestamp()
        except ValueError:
            return None

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken in the wall-clock window [t0, t1] (the timed region); the sampler itself is
        started before the warm-up because nvidia-smi needs a moment to come up."""
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(', ') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']

        def collect(windowed):
            sm, reasons = [], set()
            for r in rows:
                if len(r) < 9:
                    continue
                ts = self._epoch(r[0])
                if windowed and t0 is not None and ts is not None and not (t0 - 0.05 <= ts <= t1 + 0.05):
                    continue
                try:
                    sm.append(float(r[1]))
                    out['sm_max_mhz'] = float(r[2])
                except ValueError:
                    continue
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(n)
            return sm, reasons

        sm, reasons = collect(True)
        if not sm:      # nothing fell into the window (clock skew between nvidia-smi and this process): use every sample
            sm, reasons = collect(False)
            out['window'] = 'whole run (no sample inside the timed region)'
        if sm:
            out['sm_mhz'] = float(np.median(sm))
            out['samples'] = len(sm)
        out['reasons'] = sorted(reasons)
        return out


def make_clip(seed):
    from rib.synth import synth_flow, synth_image, synth_joints
    key = synth_image(N_KEY, H, W, seed=seed)
    joints = torch.from_numpy(synth_joints(T, H, W, seed=seed))
    flows = synth_flow(T, H, W, seed=seed)
    return key, joints, flows


# --------------------------------------------------------------------------------------------
# CPU baseline (the oracle "port"; the unmodified reference when /root/reference is importable)
# --------------------------------------------------------------------------------------------
def cpu_frames_per_s(n_frames, seed=0, warm=1):
    """Times rasterise + warp + generator + composite for `n_frames` generated frames, one at a time
    (the reference's loop is batch 1, evaluator.py:238-266), on all host cores."""
    from oracle import generator_oracle as go
    from oracle import raster_oracle as ro
    from oracle import ref_import
    from rib.arch import Arch
    from rib.config import default_gen_cfg
    from rib.synth import synth_state_dict
    torch.set_num_threads(os.cpu_count() or 1)
    arch = Arch(default_gen_cfg())
    sd = synth_state_dict(arch, seed=0)
    kind = 'port'
    ref_gen = ref_ds = None
    if ref_import.available():
        try:
            ref_gen = ref_import.make_generator()
            ref_gen.load_state_dict(sd, strict=True)
            ref_gen.eval()
            ref_ds = ref_import.make_dataset(H, W)
            kind = 'reference'
        except Exception:
            ref_gen = ref_ds = None
    key, joints, flows = make_clip(seed)
    joints = joints.numpy()

    def one(i):
        lm = [(a[0], a[1]) for a in joints[i]]
        conf = [a[2] for a in joints[i]]
        if ref_ds is not None:
            sk = ref_ds._generate_skeleton(lm, conf, H, W)
            pm = ref_ds._generate_pose_map(lm, conf, H, W)
            label = torch.cat([ref_ds.to_tensor_norm(sk)[None], torch.from_numpy(pm).float()[None]], dim=1)
        else:
            label = torch.from_numpy(ro.label(lm, conf, H, W))[None]
        prev = key[i // RATE][None]
        dain = go.warp(prev, flows[i][None])
        with torch.no_grad():
            if ref_gen is not None:
                img, mask = ref_gen(label, None, dain, prev)
            else:
                img, mask = go.generator_forward(sd, arch, label, dain, prev)
            fuse = go.composite(img, mask, dain)
            go.to_uint8(fuse)

    frames = [i for i in range(T) if i % RATE][:n_frames + warm]
    for i in frames[:warm]:
        one(i)
    t0 = time.perf_counter()
    for i in frames[warm:]:
        one(i)
    dt = time.perf_counter() - t0
    return len(frames[warm:]) / dt, kind, torch.get_num_threads()


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    n_frames = 2
    vals, kind, cores = [], 'port', 1
    for _ in range(args.warmup):
        cpu_frames_per_s(1, warm=0)
    t_all = 0.0
    for s in range(args.steps):
        fps, kind, cores = cpu_frames_per_s(n_frames, seed=s, warm=0)
        vals.append(fps)
        t_all += n_frames / fps
    value = (n_frames * args.steps) / t_all
    line = {
        'impl': 'reference', 'metric': 'rendered frames/s (raster+warp+gen+blend)', 'value': value, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_all / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                         'sample': '%d generated frames per step of the 512x512 clip, batch 1, all host threads' % n_frames},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    return {'workload': 'BASELINE configs[2]: autoregressive 2x interpolation, 65-frame clip (33 key + 32 generated) at '
                        '512x512, rasterise + flow warp + generator + mask blend; one clip per GPU per step',
            'height': H, 'width': W, 'frames_per_clip': T, 'generated_frames_per_clip': GEN_FRAMES,
            'sample_rate': RATE, 'clips_per_step': n_gpus, 'generator_batch': GEN_FRAMES,
            'l2': 'per-step working set (~20 GB of activations) exceeds the 126 MB L2; no flush needed',
            'parallelism': 'clips sharded across GPUs, no collective in the forward, NCCL all_gather of uint8 frames'}


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    import rib
    from rib.clip import ClipRenderer
    from rib.config import default_gen_cfg
    from rib.arch import Arch
    from rib.dist import gather_frames
    from rib.generator import Generator
    from rib.synth import synth_state_dict
    from rib._lib import lib
    import ctypes as C

    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU path)'
    assert lib.rib_debug_get_simt() == 0
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)

    cfg = default_gen_cfg()
    gen = Generator(cfg)
    gen.load_state_dict(synth_state_dict(Arch(cfg), seed=0), strict=True)
    gen = gen.to(dev).eval()
    renderer = ClipRenderer(gen, sample_rate=RATE)

    key_h, joints_h, flows_h = make_clip(seed=rank)
    key_h, joints_h, flows_h = key_h.pin_memory(), joints_h.pin_memory(), flows_h.pin_memory()
    key_d, joints_d, flows_d = key_h.to(dev), joints_h.to(dev), flows_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        out = renderer.render(key_d, joints_d, flows=flows_d, want_u8=True, want_fuse=False)
        if world > 1:
            gather_frames(out['u8'], world * T)
        return out

    # End-to-end pipeline: every step uploads ITS inputs from pinned host memory and downloads ITS uint8 frames,
    # all inside the timed region; uploads of step i+1 and downloads of step i-1 run on their own streams so that
    # they overlap the kernels of step i (double-buffered inputs, stream-ordered events, no host syncs in the loop).
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    in_bufs = [(torch.empty_like(key_d), torch.empty_like(joints_d), torch.empty_like(flows_d)) for _ in range(2)]
    u8_hosts = [torch.empty(T, H, W, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def run_e2e(steps):
        main = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            b = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_free[b])          # step i-2 has finished reading this buffer
                else:
                    s_in.wait_stream(main)
                for dst, src in zip(in_bufs[b], (key_h, joints_h, flows_h)):
                    dst.copy_(src, non_blocking=True)
                ev_in[b].record(s_in)

        upload(0)
        for i in range(steps):
            b = i & 1
            if i + 1 < steps:
                upload(i + 1)
            main.wait_event(ev_in[b])
            k, j, f = in_bufs[b]
            out = renderer.render(k, j, flows=f, want_u8=True, want_fuse=False)
            if world > 1:
                gather_frames(out['u8'], world * T)
            ev_free[b].record(main)
            ev_done[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                if i >= 2:
                    s_out.wait_event(ev_out[b])
                u8_hosts[b].copy_(out['u8'], non_blocking=True)
                out['u8'].record_stream(s_out)
                ev_out[b].record(s_out)
        main.wait_stream(s_out)                          # the caller holds every frame on the host
        main.wait_stream(s_in)

    def step_e2e():
        run_e2e(1)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step_resident()
        l0 = lib.rib_kernel_launch_count()
        w0 = time.time()
        ms = timed(step_resident, args.steps)
        w1 = time.time()
        launches = lib.rib_kernel_launch_count() - l0
        clocks = sampler.stop(w0, w1) if rank == 0 else None
        for _ in range(2):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_e2e(args.steps)
        e1.record()
        barrier()
        ms_t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        ms_e2e = float(ms_t.item())
        # roofline pass: the same steps with every implicit-GEMM launch bracketed by CUDA events
        lib.rib_profile_enable(1)
        barrier()
        for _ in range(args.steps):
            step_resident()
        barrier()
        conv_ms, conv_n = C.c_double(), C.c_longlong()
        lib.rib_profile_collect(C.byref(conv_ms), C.byref(conv_n))
        lib.rib_profile_enable(0)

    frames = GEN_FRAMES * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    tc_peak, hbm_peak, peak_src = peaks()
    conv_flop_per_step = CONV_FLOP_PER_FRAME * GEN_FRAMES
    # DRAM traffic of the implicit-GEMM launches: from the committed ncu launch list of this same command
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'conv_gemm_traffic.json')
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
    launches_per_step = conv_n.value / max(args.steps, 1)
    conv_ms_per_step = conv_ms.value / max(args.steps, 1)
    achieved = conv_flop_per_step / (conv_ms_per_step * 1e-3) / 1e12
    line = {
        'metric': 'rendered frames/s (raster+warp+gen+blend)', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp16' if lib.rib_act_is_fp16() else 'bf16', 'data': 'synthetic',
        'config': workload_config(world),
        'e2e': {'value': e2e_value, 'unit': 'frames/s',
                'h2d_bytes_per_step': int(key_h.numel() * 4 + joints_h.numel() * 8 + flows_h.numel() * 4),
                'd2h_bytes_per_step': int(u8_hosts[0].numel()), 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {
            'bound': 'tensor', 'kernel': 'rib::conv_gemm_kernel<EPI_STORE|EPI_SPADE|EPI_FINAL> (tcgen05 implicit GEMM)',
            'achieved': achieved, 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': achieved / tc_peak, 'peak_source': peak_src,
            'launches_per_step': launches_per_step, 'kernel_ms_per_step': conv_ms_per_step,
            'algorithmic_flop_per_step': conv_flop_per_step,
            'share_of_step': conv_ms_per_step / (ms / args.steps),
            'traffic': traffic, 'traffic_unit': 'bytes per launch (mean over the %d launches of a step)' % round(launches_per_step),
            'traffic_source': traffic_src,
            'note': 'achieved = 883.5 kFLOP/pixel x 512x512 x 32 frames / summed CUDA-event time of all conv_gemm '
                    'launches of a step (events on the launching stream, separate pass of the same steps)'},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, kind, cores = cpu_frames_per_s(3, seed=0, warm=1)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                                'sample': '3 generated frames of the same 512x512 clip after 1 warm-up frame, batch 1'}
    else:
        line['cpu_baseline'] = None
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
