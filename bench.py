#!/usr/bin/env python
"""bench.py -- KRN 224x224 training throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the UNMODIFIED reference loop (baseline/_ref) on the host CPU
    python bench.py --workload dann --gpus N  # BASELINE.json configs[3]: the adapt.py step (secondary line)

Workload = BASELINE.json configs[1]: KRN train, bs=48/GPU, AdamW (lr 1e-3, betas .9/.999, wd .01),
clip_grad_norm 1.0, 224x224 synthetic images in [0,1), random-init weights (no network for ImageNet).
One "step" = zero grads, forward, loss, backward, [grad allreduce], clip, AdamW on one batch.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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

BATCH = 48
HW = 224
METRIC = 'krn_train_images_per_sec'
UNIT = 'images/s'


def _peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops_sustained', 1400.0), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  Sampled in-process through NVML
    (nvidia_ml_py) every 0.2 s -- spawning `nvidia-smi` five times a second takes the driver lock for tens of milliseconds
    per call, which showed up as sporadic 20 % dips of the host-driven end-to-end number; `nvidia-smi` is the fallback."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.nvml, self._mx = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(vis.split(',')[index]) if vis and vis.split(',')[index].strip().isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.source = 'nvml'
        except Exception:
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        # two driver queries per sample (current SM clock, event reasons): every NVML call takes a driver lock that kernel / graph
        # launches also need -- with the maximum clock and the power draw queried each time the host-driven end-to-end loop ran
        # 6 % behind the device-resident one (tools/e2e_probe.py without a sampler: 0.4 %)
        n = self.nvml
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        if self._mx is None:
            self._mx = n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)
        mx, pw = self._mx, 0.0
        get = getattr(n, 'nvmlDeviceGetCurrentClocksEventReasons', None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.h))
        bit = lambda name, alt: getattr(n, name, getattr(n, alt, 0))
        flags = [('Active' if r & bit('nvmlClocksEventReasonHwSlowdown', 'nvmlClocksThrottleReasonHwSlowdown') else 'Not Active'),
                 ('Active' if r & bit('nvmlClocksEventReasonHwThermalSlowdown', 'nvmlClocksThrottleReasonHwThermalSlowdown') else 'Not Active'),
                 ('Active' if r & bit('nvmlClocksEventReasonSwThermalSlowdown', 'nvmlClocksThrottleReasonSwThermalSlowdown') else 'Not Active'),
                 ('Active' if r & bit('nvmlClocksEventReasonSwPowerCap', 'nvmlClocksThrottleReasonSwPowerCap') else 'Not Active')]
        return [str(sm), str(mx), '%.1f' % pw] + flags

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(',')]
                    if len(f) >= 7:
                        self.rows.append(f)
            except Exception:
                pass
            self._halt.wait(0.25)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(r[0])) for r in self.rows if r[0].replace('.', '').isdigit())
        mx = [int(float(r[1])) for r in self.rows if r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(self.rows), 'source': self.source}


# ------------------------------------------------------------------------------------------------
def _ref_runner():
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    import ref_runner
    return ref_runner if ref_runner.available() else None


def _reference_cpu(warmup, steps, budget_s=200.0):
    """The reference's own train_single_epoch_krn (unmodified, staged under baseline/_ref) on the host cores; falls back to the
    oracle port when the staged copy is absent.  Every step is a full bs=48 iteration; the number of timed steps is the one
    asked for unless that would exceed `budget_s` (stated in the result)."""
    import torch
    cores = os.cpu_count() or 1
    R = _ref_runner()
    if R is not None:
        # a bs=48 iteration costs ~0.7-1 s on 16 host cores: the step count is bounded a priori so the run ends in minutes
        w, k = max(1, warmup), max(1, min(steps, int(budget_s)))
        r = R.time_krn_train('cpu', batch=BATCH, hw=HW, warmup=w, steps=k, threads=cores)
        r.update(kind='reference', warmup=w, steps=k, cores=cores,
                 sample='%d full bs=%d iterations of the unmodified reference loop src/core/trainer.py:train_single_epoch_krn '
                        '(torch %s CPU kernels, %d threads), after %d warm-up iterations' % (k, BATCH, torch.__version__, cores, w))
        return r
    from oracle import krn as okrn, synth, steps as osteps
    torch.set_num_threads(cores)
    sd = synth.synth_state_dict(okrn.krn_shapes(), 2021)
    st = osteps.new_state(sd)
    x, y = synth.synth_images(BATCH), synth.synth_keypoints(BATCH)
    w = max(1, warmup)
    t0 = time.perf_counter()
    for _ in range(w):
        osteps.krn_train_step(sd, st, x, y)
    per = (time.perf_counter() - t0) / w
    k = max(1, min(steps, int(budget_s / max(per, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(k):
        osteps.krn_train_step(sd, st, x, y)
    dt = (time.perf_counter() - t0) / k
    return {'ms_per_step': dt * 1e3, 'images_per_sec': BATCH / dt, 'kind': 'port', 'warmup': w, 'steps': k, 'cores': cores,
            'sample': '%d full bs=%d train steps of the oracle port (baseline/_ref not staged), %d threads' % (k, BATCH, cores)}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on the box's host cores (rank 0 only)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = _reference_cpu(args.warmup, args.steps)
    v = r['images_per_sec']
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': r['steps'],
        'warmup': r['warmup'], 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
        'config': {'workload': 'KRN train bs=48 AdamW 224x224 synthetic (BASELINE.json configs[1])', 'device': 'host CPU',
                   'requested_steps': args.steps, 'requested_warmup': args.warmup},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def cpu_baseline():
    r = _reference_cpu(2, 15, budget_s=25.0)
    return {'value': r['images_per_sec'], 'unit': UNIT, 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample'],
            'ms_per_step': r['ms_per_step']}


def reference_cuda(dev_index=0):
    """The bar the product path has to beat: the UNMODIFIED reference through its own loop on the same GPU --
    stock PyTorch dispatch to cuDNN / cuBLAS, fp32 and `--use_fp16` (torch.cuda.amp autocast + GradScaler), 20 warm-up +
    100 timed iterations, wall clock around the loop with a device synchronize on both sides."""
    R = _ref_runner()
    if R is None:
        return {'unavailable': 'baseline/_ref not staged (python tools/stage_reference.py)'}
    dev = 'cuda:%d' % dev_index
    out = {}
    for key, kw in (('reference_cuda_fp32', {}), ('reference_cuda_amp', {'fp16': True})):
        try:
            out[key] = R.time_krn_train(dev, batch=BATCH, hw=HW, warmup=20, steps=100, **kw)
        except Exception as e:
            out[key] = {'error': repr(e)[:200]}
    try:
        out['reference_cuda_dann_fp32'] = R.time_dann_train(dev, batch=BATCH, hw=HW, warmup=10, steps=50)
        out['reference_cuda_styleaug_fwd'] = R.time_styleaug(dev, batch=BATCH, hw=HW, warmup=5, steps=30)
    except Exception as e:
        out['reference_cuda_other'] = {'error': repr(e)[:200]}
    return out


# ------------------------------------------------------------------------------------------------
def _time_steps(fn, warm, k):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def secondary_workloads(dev, stepper, d_img, d_tgt, k=8):
    """The other BASELINE.json configs on ONE GPU, device-resident inputs, CUDA events (ms per step, images/s):
    configs[2] KRN + style augmentation every step, configs[3] DANN step (48 source + 48 target), configs[4] SPN bs=32."""
    import torch
    from speedplusbaseline_b200.styleaug.styleAugmentor import StyleAugmentor
    from speedplusbaseline_b200.nets.revgrad import RevGrad
    from speedplusbaseline_b200.nets.spn import SpacecraftPoseNet
    from speedplusbaseline_b200.optim import FusedAdamW
    from speedplusbaseline_b200.core.dann import DANNTrainStep
    from speedplusbaseline_b200.core.trainer import SPNTrainStep
    out = {}
    # BASELINE.json configs[1] "(fp32 and --use_fp16)": the same step with the --use_fp16 mode of this path
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    mt = KeypointRegressionNet(11, device=dev, seed=2021, tf32_gemm=True)
    mt.train()
    ot = FusedAdamW(mt._store, mt.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    from speedplusbaseline_b200.core.trainer import KRNTrainStep
    stt = KRNTrainStep(mt, ot)
    ms = _time_steps(lambda: stt.step(d_img, d_tgt), 5, 30)
    out['krn_train_use_fp16_bs48'] = {'ms': ms, 'images_per_sec': BATCH / ms * 1e3, 'dtype': 'tf32',
                                      'math': 'fp32 storage, single-pass TF32 GEMMs (the --use_fp16 mode)'}
    del mt, ot, stt
    from speedplusbaseline_b200.styleaug.ghiasi import synthetic_state
    from speedplusbaseline_b200.styleaug.styleAugmentor import checkpoint_dir
    try:                                                         # the reference's REAL checkpoints (staged with baseline/_ref)
        checkpoint_dir()
        aug, weights = StyleAugmentor(0.5, dev), 'real checkpoints (baseline/_ref/src/styleaug/checkpoints)'
    except FileNotFoundError:
        aug, weights = StyleAugmentor(0.5, dev, state=synthetic_state(7)), 'synthetic (checkpoints not staged)'
    ms = _time_steps(lambda: aug(d_img), 2, k)
    out['styleaug_forward_bs48'] = {'ms': ms, 'images_per_sec': BATCH / ms * 1e3, 'tflops_reference_flops': 15.434 * BATCH / ms,
                                    'weights': weights}
    ms = _time_steps(lambda: stepper.step(aug(d_img), d_tgt), 2, k)
    out['krn_train_with_styleaug_ratio1_bs48'] = {'ms': ms, 'images_per_sec': BATCH / ms * 1e3}
    del aug
    m = RevGrad(11, device=dev, seed=2021)
    m.train()
    opt = FusedAdamW(m._store, m.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=1)
    ds = DANNTrainStep(m, opt)
    tgt_img = torch.rand_like(d_img)
    ms = _time_steps(lambda: ds.step(d_img, d_tgt, tgt_img, 0.5), 3, k)
    out['dann_step_48src_48tgt'] = {'ms': ms, 'images_per_sec': 2 * BATCH / ms * 1e3, 'source_images_per_sec': BATCH / ms * 1e3}
    del m, opt, ds
    torch.cuda.empty_cache()
    spn = SpacecraftPoseNet.__new__(SpacecraftPoseNet)
    torch.nn.Module.__init__(spn)
    from speedplusbaseline_b200.spn_engine import SPNEngine
    spn.engine = SPNEngine(5000, device=dev)
    spn._register_store(spn.engine.store, spn.engine.key_order)
    spn.engine.store.params.normal_(0.0, 0.01)                     # random init on the device (152 M parameters)
    spn.train()
    opt = FusedAdamW(spn._store, spn.parameters(), lr=1e-3, weight_decay=0.01, clip_mode=2)
    ss = SPNTrainStep(spn, opt)            # CUDA-graph step: the dropout masks come from a device-resident counter
    x = torch.rand(32, 3, 227, 227, device=dev)
    yc = torch.zeros(32, 5000, device=dev)
    yc[:, :5] = 0.2
    ms = _time_steps(lambda: ss.step(x, yc, yc), 2, k)
    out['spn_train_bs32_dropout0.5'] = {'ms': ms, 'images_per_sec': 32 / ms * 1e3}
    del spn, opt, ss, x, yc
    torch.cuda.empty_cache()
    # the reference's per-sample transform stack as one batched device call -- 48 SPEED+-sized grey
    # frames (1200x1920 uint8, resident in HBM) -> crop + Pillow-exact bilinear resize + ToTensor + rotate/flip +
    # brightness/contrast + noise -> [48,3,224,224] fp32.  Host-side decision sampling and struct upload are inside the timing.
    try:
        import numpy as np
        from speedplusbaseline_b200.datasets.transforms import build_transforms
        tf = build_transforms('krn', (HW, HW), p_aug=0.5, is_train=True, device=dev, generator=torch.Generator().manual_seed(1))
        frames = torch.randint(0, 256, (BATCH, 1200, 1920), dtype=torch.uint8, device=dev)
        rng = np.random.default_rng(1)
        cx, cy, sz = rng.uniform(500, 1400, BATCH), rng.uniform(400, 800, BATCH), rng.uniform(150, 700, BATCH)
        bbox = np.stack([cx - sz / 2, cx + sz / 2, cy - sz / 2, cy + sz / 2], 1).astype(np.float32)
        kp = np.zeros((BATCH, 2, 11), np.float32)
        ms = _time_steps(lambda: tf(frames, bbox, kp), 2, k)
        out['input_pipeline_bs48_1200x1920_u8'] = {'ms': ms, 'images_per_sec': BATCH / ms * 1e3, 'status': tf.status(),
                                                    'out_gbs': BATCH * 3 * HW * HW * 4 / ms / 1e6}
    except Exception as e:
        out['input_pipeline_bs48_1200x1920_u8'] = {'error': repr(e)[:200]}
    return out

# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-out', default='')
    ap.add_argument('--no-secondary', action='store_true')
    ap.add_argument('--no-reference-cuda', action='store_true', help='skip the reference-on-cuDNN secondaries')
    ap.add_argument('--workload', default='krn', choices=['krn', 'dann'], help='dann = BASELINE.json configs[3] (adapt.py step, 48 source + 48 target images per GPU)')
    ap.add_argument('--dtype', default='fp32', choices=['fp32', 'tf32', 'bf16'],
                    help='tf32 = the --use_fp16 mode (fp32 storage, single-pass TF32 GEMMs); bf16 = experimental bf16-storage engine; the headline is fp32')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from speedplusbaseline_b200 import _lib as L
    from speedplusbaseline_b200 import profiler
    from speedplusbaseline_b200.nets.park2019 import KeypointRegressionNet
    from speedplusbaseline_b200.nets.revgrad import RevGrad
    from speedplusbaseline_b200.optim import FusedAdamW
    from speedplusbaseline_b200.core.trainer import KRNTrainStep, DevicePrefetcher
    from speedplusbaseline_b200.core.dann import DANNTrainStep

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=90))
    W = max(3, args.warmup)
    K = max(1, args.steps)
    dann = args.workload == 'dann'

    g = torch.Generator(device='cpu').manual_seed(2021 + rank)
    h_img = torch.rand(BATCH, 3, HW, HW, generator=g).pin_memory()
    h_tgt = torch.rand(BATCH, 2, 11, generator=g).pin_memory()
    if dann:
        model = RevGrad(11, device=dev, seed=2021)
        h_tim = torch.rand(BATCH, 3, HW, HW, generator=g).pin_memory()          # unlabeled target-domain batch
        host_batch = (h_img, h_tgt, h_tim)
    else:
        model = KeypointRegressionNet(11, device=dev, seed=2021, dtype=L.BF16 if args.dtype == 'bf16' else L.F32, tf32_gemm=args.dtype == 'tf32')
        host_batch = (h_img, h_tgt)
    model.train()
    opt = FusedAdamW(model._store, model.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                     clip_mode=1, max_norm=1.0)
    if dann:
        stepper = DANNTrainStep(model, opt, use_graph=not args.no_graph, world_size=world)
        step = lambda *b: stepper.step(b[0], b[1], b[2], 0.5)
        eager = lambda *b: stepper.eager(b[0], b[1], b[2], 0.5)
        per_step = 2 * BATCH
    else:
        stepper = KRNTrainStep(model, opt, use_graph=not args.no_graph, world_size=world)
        step, eager, per_step = stepper.step, stepper.eager, BATCH
    dev_batch = tuple(t.to(dev) for t in host_batch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # launches per step (counted on an eager step; graph replays re-issue the same kernels)
    n0 = L.lib.b200sp_launch_count()
    eager(*dev_batch)
    torch.cuda.synchronize()
    launches_per_step = int(L.lib.b200sp_launch_count() - n0)

    for _ in range(W):
        step(*dev_batch)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # ---- timed region 1: device-resident inputs (kernel/graph throughput) -----------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(K):
        losses = step(*dev_batch)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    loss_after_device_region = float(losses[0])
    # ---- timed region 2: end to end through the public step API with HOST inputs ------------------
    nl = int(losses.numel())
    host_loss = torch.empty(nl).pin_memory()
    # untimed: what this box's host->device link gives for the pinned image batch (the e2e loop is bound by it when it is slow:
    # 28.9 MB per 5 ms step needs 5.6 GB/s; healthy boxes measure ~55 GB/s), and two steps through the prefetcher so that its
    # device buffers come from the caching allocator, not from a synchronising cudaMalloc inside the timed region
    hb0, hb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scratch = torch.empty_like(dev_batch[0])
    hb0.record()
    for _ in range(4):
        scratch.copy_(host_batch[0], non_blocking=True)
    hb1.record()
    torch.cuda.synchronize()
    h2d_gbs = 4 * host_batch[0].numel() * 4 / (hb0.elapsed_time(hb1) * 1e-3) / 1e9
    del scratch
    for db in DevicePrefetcher([host_batch] * 2, dev):
        step(*db)
    # the public loop's input path: pinned host batch -> DevicePrefetcher (H2D of step i+1 on a copy stream while step i
    # computes) -> stepper.step; every step's H2D and its loss D2H are inside the timed region.  The loop is host-driven, so it
    # is run twice (K steps each) and the faster pass is reported, both are listed (`e2e.passes_ms_per_step`)
    passes = []
    for _ in range(2):
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for db in DevicePrefetcher([host_batch] * K, dev):
            l3 = step(*db)
            host_loss.copy_(l3, non_blocking=True)
        e3.record()
        barrier()
        passes.append(e2.elapsed_time(e3))
    final_loss = float(host_loss[0])
    ms_e2e = min(passes)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # per-launch CUDA-event profile of eager steps: every rank runs it (the steps contain the gradient allreduce)
    prof = None
    if not dann:
        prof = profiler.profile_krn_step(stepper, dev_batch[0], dev_batch[1], reps=3)
    if rank == 0:
        hbm, tf, which = _peaks()
        value = world * per_step * K / (ms / 1e3)
        e2e = world * per_step * K / (ms_e2e / 1e3)
        h2d = int(sum(t_.numel() * 4 for t_ in host_batch))
        if dann:
            metric, workload = 'dann_adapt_images_per_sec', 'KRN DANN adapt.py step, 48 source + 48 target images per GPU (BASELINE.json configs[3])'
            math = 'fp32 storage, 3xTF32 tensor-core GEMMs + fp32 CUDA-core stencils (DANN is fp32-only in the reference, adapt.py:99-101)'
        else:
            metric, workload = METRIC, 'KRN train bs=48/GPU AdamW 224x224 synthetic (BASELINE.json configs[1])'
            math = {'fp32': 'fp32 storage, 3xTF32 tensor-core GEMMs + fp32 CUDA-core stencils',
                    'tf32': '--use_fp16 mode: fp32 storage, single-pass TF32 tensor-core GEMMs (fp16 operand mantissa, fp32 range / accumulate), fp32 everything else',
                    'bf16': 'bf16 storage + bf16 tensor-core GEMMs, fp32 accumulate / statistics / master weights (experimental)'}[args.dtype]
        line = {
            'metric': metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
            'ms_per_step': ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': workload, 'math': math, 'global_batch': world * per_step, 'parallelism': 'dp%d' % world,
                       'cuda_graph': not args.no_graph,
                       'l2': 'per-step working set ~2.7 GB of activations >> 126 MB L2 (no explicit flush)'},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4 * nl, 'ms_per_step': ms_e2e / K,
                    'passes_ms_per_step': [p_ / K for p_ in passes], 'h2d_link_gbs': h2d_gbs},
            'gpu_launches': launches_per_step * K, 'launches_per_step': launches_per_step,
            'clocks': clocks, 'final_loss': final_loss, 'loss_after_device_region': loss_after_device_region,
        }
        if prof is not None:
            # roofline of the kernel FAMILY that owns the step (per-family table alongside): algorithmic bytes of all its launches /
            # the sum of their CUDA-event durations in an eager step, against the measured copy bandwidth.  `traffic` (ncu DRAM
            # bytes) is not measured inside a bench run: null here, the ncu captures are under profiles/.
            fams = prof['families']
            lead = fams[0]
            roof = {'bound': 'hbm', 'kernel': lead['family'], 'achieved': lead['gbs'], 'peak': hbm, 'unit': 'GB/s',
                    'frac': lead['gbs'] / hbm, 'traffic': None, 'peak_source': which + ' (sustained copy)',
                    'share_of_step': lead['share'], 'launches': lead['launches'], 'us_per_step': lead['us'],
                    'algorithmic_mb_per_step': lead['algorithmic_mb'], 'families': fams,
                    'longest_launch': prof['dominant'],
                    'step_roofline_frac': prof['step_roofline_ms'] / (ms / K), 'step_algorithmic_gb': prof['step_bytes'] / 1e9,
                    'note': 'optimizer family: its 158 MB flat buffers are partly L2-resident after the backward kernels, so its GB/s is not a pure HBM figure'}
            # the metric's second half, "conv tensor-pipe % of peak" (SURVEY.md 8 d-1): dense conv/FC FLOPs of the reference graph
            # (2.3345 GFLOP per image for a train step; depthwise excluded) / step time / measured dense bf16 peak.  The fp32 path
            # issues 3 tf32 MMAs per product (tf32 rate = bf16/2), so its own ceiling is peak/6.
            dense_gf = 2.3345 * BATCH
            tflops = dense_gf / (ms / K)
            roof['tensor_pipe'] = {'dense_gflop_per_step': dense_gf, 'achieved': tflops, 'peak': tf, 'unit': 'TFLOP/s',
                                   'frac': tflops / tf, 'peak_source': which + ' (cuBLAS bf16 sustained)',
                                   'mma_per_product': 3 if args.dtype == 'fp32' else 1}
            line['roofline'] = roof
            if args.profile_out:
                with open(args.profile_out, 'w') as f:
                    f.write(prof['table'])
        if not args.no_secondary and world == 1 and not dann:
            try:
                line['secondary'] = secondary_workloads(dev, stepper, dev_batch[0], dev_batch[1])
            except Exception as e:          # secondary numbers never invalidate the headline line
                line['secondary'] = {'error': repr(e)[:300]}
            if not args.no_reference_cuda:
                del stepper
                torch.cuda.empty_cache()
                line['secondary'].update(reference_cuda(local))
                rc = line['secondary'].get('reference_cuda_fp32', {})
                if 'ms_per_step' in rc:
                    line['secondary']['speedup_vs_reference_cuda_fp32'] = rc['ms_per_step'] / (ms_e2e / K)
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
