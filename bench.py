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


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner did, from the C side),
    so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, 'w')


def bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned host buffers are allocated
    (first touch puts them on that NUMA node): with 8 ranks every upload / download then crosses its own socket's PCIe
    root instead of one node's memory controllers.  Best effort: any failure leaves the affinity alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [w * 64 + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return 0


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


def to_u8_frames(x):
    """fp32 [N,3,H,W] in [-1,1] -> uint8 [N,H,W,3] (rounded): what a decoded PNG key frame looks like."""
    return ((x * 0.5 + 0.5).clamp(0, 1) * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def make_clip_lean(seed, n_key=N_KEY, rate=RATE):
    """The upload of one clip as a pipeline would hold it on the host: uint8 HWC key frames (decoded images), joints
    for every frame (19x3 doubles each) and flows for the GENERATED frames only, in frame order."""
    from rib.synth import synth_flow, synth_image, synth_joints
    t = (n_key - 1) * rate + 1
    key_u8 = to_u8_frames(synth_image(n_key, H, W, seed=seed))
    joints = torch.from_numpy(synth_joints(t, H, W, seed=seed))
    flows = synth_flow(t, H, W, seed=seed)
    gen_rows = [i for i in range(t) if i % rate]
    # flow fields travel as float16 (|flow| <= 8 px: 2^-7 px resolution at most; rib.warp converts exactly on load)
    return key_u8, joints, flows[gen_rows].contiguous().half()


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


def run_reference(args, rank, world, out):
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
    print(json.dumps(line), file=out)
    out.flush()


def workload_config(n_gpus):
    return {'workload': 'BASELINE configs[2]: autoregressive 2x interpolation, 65-frame clip (33 key + 32 generated) at '
                        '512x512, rasterise + flow warp + generator + mask blend; one clip per GPU per step',
            'height': H, 'width': W, 'frames_per_clip': T, 'generated_frames_per_clip': GEN_FRAMES,
            'sample_rate': RATE, 'clips_per_step': n_gpus, 'generator_batch': GEN_FRAMES,
            'inputs': 'uint8 key frames (decoded images), float64 joints, float16 flows of the generated frames',
            'l2': 'per-step working set (~20 GB of activations) exceeds the 126 MB L2; no flush needed',
            'parallelism': 'clips sharded across GPUs, no collective in the forward, asynchronous NCCL gather of the '
                           'uint8 frames onto rank 0 (grouped send/recv, overlaps the next clip)'}


# --------------------------------------------------------------------------------------------
# Same-box bar: the reference generator's dataflow (the functional fp32
# restatement in oracle/, which test_oracle_vs_reference pins bit-for-bit to the reference nn.Module) executed by
# stock PyTorch / cuDNN on the same B200.  A baseline leg: nothing here is on the product path.
# --------------------------------------------------------------------------------------------
def aten_baseline(dev, gen, batch, iters=3):
    from oracle import generator_oracle as go
    from rib.arch import Arch
    from rib.config import default_gen_cfg
    from rib.synth import synth_image, synth_state_dict
    arch = Arch(default_gen_cfg())
    sd = {k: v.to(dev) for k, v in synth_state_dict(arch, seed=0, power_iters=5).items()}
    label = torch.rand(batch, 22, H, W, device=dev) * 2 - 1
    fake, prev = synth_image(batch, H, W, seed=1).to(dev), synth_image(batch, H, W, seed=2).to(dev)
    out = {'batch': batch, 'height': H, 'width': W, 'iters': iters,
           'what': 'oracle/generator_oracle.generator_forward (== reference Generator.forward) run by ATen/cuDNN on cuda'}

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        with torch.no_grad():
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
            out['generator_ms_fp32'] = timed(lambda: go.generator_forward(sd, arch, label, fake, prev))
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
            out['generator_ms_tf32'] = timed(lambda: go.generator_forward(sd, arch, label, fake, prev))
            cl = torch.channels_last
            sd_cl = {k: (v.contiguous(memory_format=cl) if v.dim() == 4 else v) for k, v in sd.items()}
            lab_cl, fake_cl, prev_cl = (t.contiguous(memory_format=cl) for t in (label, fake, prev))

            def bf16():
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    go.generator_forward(sd_cl, arch, lab_cl, fake_cl, prev_cl)
            out['generator_ms_bf16_autocast_channels_last'] = timed(bf16)
            out['generator_ms_ours'] = timed(lambda: gen(label, None, fake, prev))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    best = min(out['generator_ms_fp32'], out['generator_ms_tf32'], out['generator_ms_bf16_autocast_channels_last'])
    out['generator_frames_per_s'] = {k[len('generator_ms_'):]: batch / (v * 1e-3) for k, v in out.items()
                                     if k.startswith('generator_ms_')}
    out['ours_vs_best_aten'] = best / out['generator_ms_ours']
    out['ours_vs_fp32_aten'] = out['generator_ms_fp32'] / out['generator_ms_ours']
    del sd, sd_cl
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------------------------
# BASELINE configs[3]: 4x interpolation of 256 synthetic clips sharded across the ranks (strong scaling), end to end:
# every clip is uploaded from pinned host memory, rendered (3 dependent AR passes of batch 16), downloaded, and
# gathered onto rank 0 asynchronously.
# --------------------------------------------------------------------------------------------
C4_CLIPS, C4_KEY, C4_RATE = 256, 17, 4
C4_T = (C4_KEY - 1) * C4_RATE + 1


def run_c4(gen, dev, rank, world, n_clips=C4_CLIPS):
    import torch.distributed as dist
    from rib.clip import ClipRenderer
    from rib.dist import gather_frames, shard_range
    assert n_clips % world == 0, 'configs[3] shards 256 clips evenly over 1/2/4/8 GPUs'
    renderer = ClipRenderer(gen, sample_rate=C4_RATE)
    lo, hi = shard_range(n_clips, rank, world)
    mine = hi - lo
    # Clips are independent: two of them form one generator batch per AR step (2 x 16 intervals = batch 32, the
    # efficiency of the 2x clip); an odd remainder is rendered alone.
    group = 2 if mine % 2 == 0 else 1
    steps_total = mine // group
    clips = [make_clip_lean(1000 + rank * 2 + j, C4_KEY, C4_RATE) for j in range(2)]
    pool = tuple(torch.stack([clips[j % 2][i] for j in range(group)]).pin_memory() for i in range(3))   # [group, ...] per input
    dbuf = [tuple(torch.empty_like(t, device=dev) for t in pool) for _ in range(2)]
    hosts = [torch.empty(group * C4_T, H, W, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
    big = None
    if world > 1 and rank == 0:     # [step][rank][clip of the step][frame]: 13 GB of uint8 frames for 256 clips
        big = torch.empty(steps_total, world * group * C4_T, H, W, 3, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()

    def run(count):
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        pend = []

        def upload(i):
            b = i & 1
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(ev_free[b])
                else:
                    s_in.wait_stream(main)
                for dst, src in zip(dbuf[b], pool):
                    dst.copy_(src, non_blocking=True)
                ev_in[b].record(s_in)

        upload(0)
        for i in range(count):
            b = i & 1
            if i + 1 < count:
                upload(i + 1)
            main.wait_event(ev_in[b])
            k, j, f = dbuf[b]
            u8 = renderer.render_clips(k, j, flows=f).view(group * C4_T, H, W, 3)
            ev_free[b].record(main)
            if world > 1:
                pend.append(gather_frames(u8, world * group * C4_T, out=big[i] if rank == 0 else None, async_op=True))
            ev_done[b].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_done[b])
                if i >= 2:
                    s_out.wait_event(ev_out[b])
                hosts[b].copy_(u8, non_blocking=True)
                u8.record_stream(s_out)
                ev_out[b].record(s_out)
        for p in pend:
            p.wait()
        main.wait_stream(s_out)
        main.wait_stream(s_in)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run(min(2, steps_total))           # warm-up: builds the plan for this batch, touches every buffer
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps_total)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    gen_frames = n_clips * (C4_T - C4_KEY)
    del big
    return {'workload': 'BASELINE configs[3]: 4x interpolation of %d synthetic 65-frame clips (17 key + 48 generated, 3 '
                        'dependent AR passes) at 512x512, clips sharded contiguously over the ranks' % n_clips,
            'value': gen_frames / (ms * 1e-3), 'unit': 'frames/s', 'scaling': 'strong', 'n_gpus': world, 'clips': n_clips,
            'clips_per_generator_batch': group, 'generator_batch': group * (C4_KEY - 1),
            'generated_frames': gen_frames, 'ms_total': ms,
            'includes': 'per-clip H2D from pinned memory (uint8 keys, joints, flows of generated frames), per-clip D2H '
                        'of the uint8 frames, asynchronous NCCL gather of every clip onto rank 0',
            'h2d_bytes_per_clip': int(sum(t.numel() * t.element_size() for t in pool) // group),
            'd2h_bytes_per_clip': int(hosts[0].numel() // group)}


# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-aten', action='store_true', help='skip the same-box ATen/cuDNN generator timing')
    ap.add_argument('--no-c4', action='store_true', help='skip the configs[3] (4x, 256 clips, strong scaling) pass')
    ap.add_argument('--c4-clips', type=int, default=C4_CLIPS)
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    out = protect_stdout()

    if args.impl == 'reference':
        run_reference(args, rank, world, out)
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
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else 0
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)

    cfg = default_gen_cfg()
    gen = Generator(cfg)
    gen.load_state_dict(synth_state_dict(Arch(cfg), seed=0), strict=True)
    gen = gen.to(dev).eval()
    renderer = ClipRenderer(gen, sample_rate=RATE)

    key_h, joints_h, flows_h = (t.pin_memory() for t in make_clip_lean(rank))
    key_d, joints_d, flows_d = key_h.to(dev), joints_h.to(dev), flows_h.to(dev)
    gather_out = None
    if world > 1 and rank == 0:
        gather_out = [torch.empty(world * T, H, W, 3, dtype=torch.uint8, device=dev) for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    in_bufs = [(torch.empty_like(key_d), torch.empty_like(joints_d), torch.empty_like(flows_d)) for _ in range(2)]
    u8_hosts = [torch.empty(T, H, W, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]

    def run_steps(steps, e2e):
        """`steps` clips back to back.  Resident mode: inputs are already in HBM, outputs stay there.  End-to-end mode:
        every step uploads ITS inputs from pinned host memory and downloads ITS uint8 frames inside the timed region;
        uploads of step i+1 and downloads of step i-1 run on their own streams so that they overlap the kernels of step
        i (double-buffered, stream-ordered events, no host syncs in the loop).  With several ranks every clip's frames
        are gathered onto rank 0 by an asynchronous grouped send/recv that overlaps the next clip."""
        main = torch.cuda.current_stream()
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]
        pend = []

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

        if e2e:
            upload(0)
        for i in range(steps):
            b = i & 1
            if e2e:
                if i + 1 < steps:
                    upload(i + 1)
                main.wait_event(ev_in[b])
                k, j, f = in_bufs[b]
            else:
                k, j, f = key_d, joints_d, flows_d
            out = renderer.render(k, j, flows=f, want_u8=True, want_fuse=False)
            if e2e:
                ev_free[b].record(main)
            if world > 1:
                if i >= 2:
                    pend[i - 2].wait()                   # its receive buffer is re-used now
                pend.append(gather_frames(out['u8'], world * T, out=gather_out[b] if rank == 0 else None, async_op=True))
            if e2e:
                ev_done[b].record(main)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[b])
                    if i >= 2:
                        s_out.wait_event(ev_out[b])
                    u8_hosts[b].copy_(out['u8'], non_blocking=True)
                    out['u8'].record_stream(s_out)
                    ev_out[b].record(s_out)
        for p in pend[-2:]:
            p.wait()
        if e2e:
            main.wait_stream(s_out)                      # the caller holds every frame on the host
            main.wait_stream(s_in)

    def timed(steps, e2e):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(steps, e2e)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def reproducibility_check():
        """SURVEY.md §8e "Check": N-GPU output identical to 1-GPU output.  Every rank renders the SAME clip (seed 4242)
        twice and the uint8 frames are reduced to a 64-bit checksum; the checksums of all ranks (and of the two runs)
        must agree bit for bit - instance-norm statistics are accumulated with order-independent integer atomics and
        every rank runs the shipped tuning table, so a clip's frames do not depend on which GPU of the job renders it."""
        k, j, f = (t.to(dev) for t in make_clip_lean(4242))
        sums = []
        for _ in range(2):
            u8 = renderer.render(k, j, flows=f, want_u8=True, want_fuse=False)['u8']
            w = torch.arange(1, u8.numel() + 1, device=dev, dtype=torch.int64) % 65521
            sums.append(int((u8.view(-1).to(torch.int64) * w).sum().item()))
        mine = torch.tensor(sums, device=dev, dtype=torch.int64)
        allr = [torch.empty_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, mine)
        else:
            allr = [mine]
        vals = sorted({int(v) for t in allr for v in t.tolist()})
        return {'clip_checksum': vals[0], 'ranks': world, 'runs_per_rank': 2, 'bit_identical': len(vals) == 1}

    with torch.no_grad():
        repro = reproducibility_check()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        run_steps(args.warmup, False)
        l0 = lib.rib_kernel_launch_count()
        w0 = time.time()
        ms = timed(args.steps, False)
        w1 = time.time()
        launches = lib.rib_kernel_launch_count() - l0
        clocks = sampler.stop(w0, w1) if rank == 0 else None
        run_steps(2, True)
        ms_e2e = timed(args.steps, True)
        # roofline pass: the same steps with every implicit-GEMM launch bracketed by CUDA events
        lib.rib_profile_enable(1)
        barrier()
        run_steps(args.steps, False)
        barrier()
        conv_ms, conv_n = C.c_double(), C.c_longlong()
        lib.rib_profile_collect(C.byref(conv_ms), C.byref(conv_n))
        lib.rib_profile_enable(0)
        aten = None
        if rank == 0 and world == 1 and not args.no_aten:
            aten = aten_baseline(dev, gen, GEN_FRAMES)
        c4 = None
        if not args.no_c4:
            c4 = run_c4(gen, dev, rank, world, args.c4_clips)

    frames = GEN_FRAMES * world * args.steps
    value = frames / (ms * 1e-3)
    e2e_value = frames / (ms_e2e * 1e-3)
    tc_peak, hbm_peak, peak_src = peaks()
    conv_flop_per_step = CONV_FLOP_PER_FRAME * GEN_FRAMES
    launches_per_step = conv_n.value / max(args.steps, 1)
    conv_ms_per_step = conv_ms.value / max(args.steps, 1)
    # DRAM traffic of the implicit-GEMM launches: from the committed ncu launch list of this same command.  The file
    # records how many conv launches a step had when it was captured; a different count means the kernels changed
    # since, and the figure is withheld rather than quoted stale.
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, 'profiles', 'conv_gemm_traffic.json')
    if os.path.isfile(tpath):
        tj = json.load(open(tpath))
        if round(tj.get('launches_per_step', -1)) == round(launches_per_step):
            traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
        else:
            traffic_src = 'withheld: %s was captured with %s conv launches per step, this run has %d' % (
                os.path.basename(tpath), tj.get('launches_per_step'), round(launches_per_step))
    achieved = conv_flop_per_step / (conv_ms_per_step * 1e-3) / 1e12
    h2d = int(sum(t.numel() * t.element_size() for t in (key_h, joints_h, flows_h)))
    line = {
        'metric': 'rendered frames/s (raster+warp+gen+blend)', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'fp16' if lib.rib_act_is_fp16() else 'bf16', 'data': 'synthetic',
        'config': workload_config(world),
        'e2e': {'value': e2e_value, 'unit': 'frames/s', 'h2d_bytes_per_step': h2d * world,
                'd2h_bytes_per_step': int(u8_hosts[0].numel()) * world, 'ms_per_step': ms_e2e / args.steps,
                'note': 'bytes are summed over the %d rank(s); every rank moves its own clip' % world,
                'host_affinity': ('rank pinned to the %d CPUs local to its GPU (NVML)' % numa_cpus) if numa_cpus else 'unchanged'},
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
        'aten_baseline': aten,
        'c4': c4,
        'reproducibility': repro,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, kind, cores = cpu_frames_per_s(3, seed=0, warm=1)
        line['cpu_baseline'] = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                                'sample': '3 generated frames of the same 512x512 clip after 1 warm-up frame, batch 1'}
    else:
        line['cpu_baseline'] = None
    if rank == 0:
        print(json.dumps(line), file=out)
        out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
