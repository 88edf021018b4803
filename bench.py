#!/usr/bin/env python
"""bench.py -- 12-s clips/sec of the piano-a2s training hot path (VQT -> ConvStack -> BiGRU -> hierarchical decoder ->
4xNLL -> backward -> clip + Adadelta) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                      (N=1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                         (N>1, NCCL)
    python bench.py --impl reference ...                               (the reference's own CPU path: oracle/_ref, else the oracle port)
    python bench.py --impl reference-gpu ...                           (same-box GPU baseline: the unmodified reference module on stock
                                                                        PyTorch / cuDNN / cuBLAS on the B200; builder-run, profiles/)
    python bench.py --workload infer [--impl reference]                (BASELINE configs[4] / [0]: batched greedy decode; not the headline)
    python bench.py --workload finetune                                (configs[3]: finetune.yaml, ragged U[4 s,12 s] clips, tf 0.6, bf16)
    python bench.py --precision bf16                                   (configs[2]: bf16 contractions; default bf16x3 = fp32-accurate split)

Workload = BASELINE.json configs[1]: pretrain.yaml model (random init, seed 1234), synthetic 12-s clips, batch 16 per
GPU, fp32, teacher forcing 0.7, targets U[40,80)/U[20,50) tokens per bar.  Weak scaling: per-GPU batch is
fixed, `value` = clips all ranks processed / max-over-ranks device time.  Prints ONE JSON line on rank 0.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(max_length=(398, 189))            # everything else is the constructor default == pretrain.yaml:84-96
N_SAMPLES = 192000
TF_RATIO = 0.7
METRIC = "clips_per_sec_fwd_bwd"
UNIT = "clips/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="train", choices=["train", "finetune", "infer"],
                    help="train = BASELINE configs[1] (the headline metric); finetune = configs[3] (finetune.yaml: ragged real-audio-shaped "
                         "clips, teacher forcing 0.6); infer = configs[4], batched greedy decode (evaluate path)")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16x3", "bf16"],
                    help="contraction precision of the training step: bf16x3 (default for train: fp32 operands as two bf16 pieces, fp32 "
                         "TMEM accumulation, 1e-4 parity) or bf16 (BASELINE configs[2..3]; default for finetune)")
    ap.add_argument("--also-steps", type=int, default=5, help="steps of the secondary bf16 measurement in the default headline line (0 = off)")
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU per step (default 16 for train, 32 for infer)")
    ap.add_argument("--cpu-clips", type=int, default=1, help="clips per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 32 if a.workload == "infer" else 16
    if a.precision is None:
        a.precision = "bf16" if a.workload == "finetune" else "bf16x3"
    return a


PRECISION_NOTE = {
    "bf16x3": "fp32 tensors; contractions on tcgen05 with every fp32 operand split into two bf16 pieces (hi*hi + hi*lo + lo*hi, fp32 TMEM "
              "accumulation): fp32-level results (1e-4 parity with the fp32 reference)",
    "bf16": "fp32 master tensors; contractions on tcgen05 with single bf16 operands, fp32 accumulation (BASELINE configs[2..3])",
    "fp32": "exact-fp32 FFMA kernels (the eval / greedy-decode mode)"}
DTYPE = {"bf16x3": "f32 (bf16x3 split operands, fp32 accumulate)", "bf16": "bf16", "fp32": "f32"}


def tf_ratio(workload):
    return 0.6 if workload == "finetune" else TF_RATIO          # finetune.yaml: constant 0.6; pretrain.yaml:41: 0.7


def workload_name(batch, workload="train", precision="bf16x3"):
    if workload == "finetune":
        return (f"BASELINE configs[3]: finetune.yaml model (same architecture as pretrain.yaml), synthetic real-audio-shaped clips: lengths "
                f"U[4 s,12 s] (asap.py:101), peak-normalised, frames beyond the clip zero-padded to 1201 (asap.py:345-349), batch {batch}/GPU, "
                f"one fwd+bwd training step + clip + Adadelta, {precision}, teacher_forcing 0.6")
    return (f"pretrain.yaml model, synthetic 12-s clips ({N_SAMPLES} samples @16 kHz -> 1201x480 VQT), batch {batch}/GPU, "
            f"one fwd+bwd training step + clip + Adadelta, {precision}, teacher_forcing {TF_RATIO}")


def make_ragged_audio(B, seed):
    """configs[3]: clip lengths U[4 s, 12 s] (the ASAP duration filter, asap.py:100-102), peak-normalised (asap.py:86), zero-padded to the
    12-s buffer -> (audio (B, N_SAMPLES) float32, n_samples (B,) int64)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(4 * 16000, 12 * 16000 + 1, (B,), generator=g)
    a = torch.clamp(0.25 * torch.randn(B, N_SAMPLES, generator=g), -1.0, 1.0)
    for b in range(B):
        a[b, int(n[b]):] = 0.
        a[b] /= a[b].abs().max()
    return a, n


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, c in self.rows:
            if len(c) < 9:
                continue
            try:
                mx = max(mx, float(c[2]))
                if t0 <= ts <= t1 + 0.25:
                    sm.append(float(c[1]))
                    for n, v in zip(names, c[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference arms
def _reference_module():
    """(reference `models` module or None, kind): the unmodified reference compiled into oracle/_ref by oracle/build_ref.py, else None
    (-> the oracle port)."""
    from oracle import build_ref
    if build_ref.available():
        return build_ref.load(), "reference"
    return None, "port"


def _synthetic_inputs(n_clips, workload, device="cpu", with_vqt=True):
    """Spectrogram factory + targets of the workload.  The VQT stage of the CPU arms is the float64 numpy restatement of librosa.vqt
    (librosa / soxr are not installed; the reference computes it offline on the CPU, utilities.py:240-254); the GPU-baseline arm is fed
    cached spectrograms like the reference's own training loop (datasets/syn.py:99-100)."""
    import numpy as np
    import torch
    from oracle import vqt_oracle as VO
    from piano_a2s_b200.synthetic import make_audio, make_ground_truth
    gt = make_ground_truth(n_clips, 5, 398, 189, seed=1234)
    if workload == "finetune":
        audio, n = make_ragged_audio(n_clips, seed=1234)
        clips = [audio[b, :int(n[b])].numpy() for b in range(n_clips)]
    else:
        clips = list(make_audio(n_clips, N_SAMPLES, seed=1234).numpy())

    def spectrogram():
        out = torch.zeros(n_clips, 1, 1201, 480)
        for b, a in enumerate(clips):
            v = torch.from_numpy(VO.get_vqt(a))
            out[b, 0, :v.shape[0]] = v
        return out
    if not with_vqt:
        cached = spectrogram().to(device)
        return (lambda: cached), [t.to(device) for t in gt]
    return spectrogram, gt


def reference_step_factory(n_clips, workload="train", device="cpu", with_vqt=True):
    """One bounded sample of the training workload through the REFERENCE's own code when oracle/_ref exists: its `ScoreTranscription`
    (train mode, python-`random` teacher forcing), the 4 NLL losses of pretrain.py:56-93, backward, `clip_grad_norm_(5.0)` and
    `torch.optim.Adadelta(lr=1, rho=.95, eps=1e-8)` (pretrain.yaml:44-47) -- on `device` ("cpu": all host threads; "cuda": stock
    PyTorch / cuDNN / cuBLAS, the same-box GPU baseline).  Without oracle/_ref: the oracle port (CPU only).  -> (step, threads, kind)."""
    import torch
    import models
    from oracle import a2s_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ref, kind = _reference_module()
    spectrogram, gt = _synthetic_inputs(n_clips, workload, device, with_vqt)
    tf = tf_ratio(workload)
    dev = torch.device(device)
    torch.manual_seed(1234)
    if ref is not None:
        model = ref.ScoreTranscription(**CFG).to(dev).train()
        opt = torch.optim.Adadelta(model.parameters(), lr=1.0, rho=0.95, eps=1e-8)

        def step():
            outs = model(spectrogram=spectrogram().to(dev), inference=False, ground_truth=gt, teacher_forcing_ratio=tf, device=dev)
            loss = O.training_loss(outs, gt)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
            opt.step()
            opt.zero_grad()
            return float(loss.detach())          # the reference reads its losses back every step (pretrain.py:90-93)
        return step, torch.get_num_threads(), kind
    if dev.type != "cpu":
        raise RuntimeError("the GPU baseline needs oracle/_ref (python -m oracle.build_ref where /root/reference exists)")
    sd = {k: v.clone() for k, v in models.ScoreTranscription(**CFG).state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd.items() if v.dtype == torch.float32 and "running" not in k}
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in params.items()}

    def step():
        outs = O.score_transcription(sd, spectrogram(), CFG, False, gt, tf, True)
        loss = O.training_loss(outs, gt)
        grads = dict(zip(params.keys(), torch.autograd.grad(loss, list(params.values()), allow_unused=True)))
        grads = {k: (g if g is not None else torch.zeros_like(params[k])) for k, g in grads.items()}
        with torch.no_grad():
            O.adadelta_step(params, grads, state)
        return float(loss.detach())
    return step, torch.get_num_threads(), kind


def reference_infer_factory(n_clips, device="cpu", with_vqt=True):
    """BASELINE configs[0] / [4]: greedy hierarchical decode (eval mode, no_grad) of `n_clips` synthetic 12-s clips through the reference
    module (oracle/_ref) or the oracle port, + argmax / unpad token lists on the host.  -> (step, threads, kind)."""
    import torch
    import models
    from oracle import a2s_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ref, kind = _reference_module()
    spectrogram, _ = _synthetic_inputs(n_clips, "train", device, with_vqt)
    dev = torch.device(device)
    torch.manual_seed(1234)
    if ref is not None:
        model = ref.ScoreTranscription(**CFG).to(dev).eval()

        def step():
            with torch.no_grad():
                outs = model(spectrogram=spectrogram().to(dev), inference=True, ground_truth=None, teacher_forcing_ratio=0., device=dev)
                return O.greedy_tokens([o.cpu() for o in outs])
        return step, torch.get_num_threads(), kind
    if dev.type != "cpu":
        raise RuntimeError("the GPU baseline needs oracle/_ref (python -m oracle.build_ref where /root/reference exists)")
    sd = {k: v.clone() for k, v in models.ScoreTranscription(**CFG).state_dict().items()}

    def step():
        with torch.no_grad():
            return O.greedy_tokens(O.score_transcription(sd, spectrogram(), CFG))
    return step, torch.get_num_threads(), kind


def _kind_text(kind):
    return ("the unmodified reference module (oracle/_ref, compiled from /root/reference)" if kind == "reference"
            else "oracle port of the reference")


def run_reference(args):
    """--impl reference: the reference's CPU path on the box's host cores.  --impl reference-gpu: the same module on the B200 with stock
    PyTorch kernels (same-box GPU baseline, SURVEY 8d last line)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gpu = args.impl == "reference-gpu"
    device = "cuda" if gpu else "cpu"
    n = args.batch if gpu else args.cpu_clips
    infer = args.workload == "infer"
    if infer:
        step, threads, kind = reference_infer_factory(n, device, with_vqt=not gpu)
        metric = "clips_per_sec_greedy_decode"
        workload = ("BASELINE configs[4]: batched greedy hierarchical decode, eval mode" if gpu else
                    "BASELINE configs[0]: single synthetic 12-s clip, random-init model, VQT + encoder + greedy hierarchical decode on CPU")
        what = "greedy decode of all 5 x (398 + 189) note steps"
    else:
        step, threads, kind = reference_step_factory(n, args.workload, device, with_vqt=not gpu)
        metric = METRIC
        workload = workload_name(args.batch, args.workload, "fp32" if not gpu else "fp32 (stock PyTorch, TF32 off)")
        what = "one fwd+bwd+clip+Adadelta step of the same workload"
    sync = torch.cuda.synchronize if gpu else (lambda: None)
    warm = max(min(args.warmup, 1), 1 if gpu else 0)
    for _ in range(warm):
        step()
    sync()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    sync()
    dt = time.time() - t0
    v = n * args.steps / dt
    where = "1 x B200, stock PyTorch (cuDNN / cuBLAS / ATen) kernels, cached spectrograms on the device" if gpu else \
            f"{threads} host threads, torch CPU fp32 + float64 numpy VQT"
    sample = f"{n} clip(s)/step, {what}, {_kind_text(kind)}, {where}"
    line = {
        "impl": args.impl, "metric": metric, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "sample": sample, "clips_per_step": n},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if gpu:
        line["gpu_baseline"] = {"value": v, "unit": UNIT, "kind": kind, "sample": sample}
    else:
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import models
    from piano_a2s_b200 import ops, train
    from piano_a2s_b200._lib import lib
    from piano_a2s_b200.synthetic import executed_steps, make_audio, make_ground_truth
    from piano_a2s_b200.vqt import VQT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    finetune = args.workload == "finetune"
    TF = tf_ratio(args.workload)
    ops.set_precision(train=args.precision)
    torch.manual_seed(1234)
    model = models.ScoreTranscription(**CFG).to(dev).train()
    model.convstack.sync_batchnorm = world > 1          # speechbrain converts BatchNorm -> SyncBatchNorm under DDP
    opt = train.FlatAdadelta(model)
    vqt = VQT().to(dev)
    # Pre-filled allocator pools (main stream + the decoder's two staff streams).  The launch thread runs several steps ahead of the
    # device, so buffers that were used on the side streams cannot be recycled until those streams have caught up, and the allocator
    # would otherwise meet the shortfall with cudaMalloc calls in the middle of a step, which can wait for the device to drain (seen
    # as 100-150 ms outlier steps).  A training script does the same once at start-up (train.reserve_memory).
    train.reserve_memory(float(os.environ.get("PA2S_RESERVE_GB", "16")), dev, streams=model.decoder.streams(), stream_gigabytes=6.0)
    if finetune:
        audio_h, n_h = make_ragged_audio(B, seed=1234 + rank)
        audio_h, n_h = audio_h.pin_memory(), n_h.pin_memory()
    else:
        audio_h, n_h = make_audio(B, N_SAMPLES, seed=1234 + rank).pin_memory(), None
    gt_h = [t.pin_memory() for t in make_ground_truth(B, 5, 398, 189, seed=1234 + rank)]
    S = executed_steps(gt_h)
    audio_d = audio_h.to(dev)
    n_d = n_h.to(dev) if finetune else None
    gt_d = train.targets_to_device(gt_h, dev)

    def step_device():
        spec = vqt(audio_d, n_d).unsqueeze(1)
        return train.fit_batch(model, opt, spec, gt_d, TF)

    def step_e2e():
        a = audio_h.to(dev, non_blocking=True)
        n = n_h.to(dev, non_blocking=True) if finetune else None
        g = train.targets_to_device(gt_h, dev)
        spec = vqt(a, n).unsqueeze(1)
        return train.fit_batch(model, opt, spec, g, TF).item()          # D2H read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms, per_step = [], []

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        marks, hmarks = [e0], [w0]
        dbg = os.environ.get("PA2S_BENCH_DEBUG") == "1"
        for _ in range(k):
            fn()
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
            hmarks.append(time.time())
            if dbg:
                st = torch.cuda.memory_stats(dev)
                print("dbg step host_ms %.1f gc %s seg_alloc %d seg_free %d reserved %.2f GB active %.2f GB retries %d" % (
                    1e3 * (hmarks[-1] - hmarks[-2]), gc.get_count(), st.get("segment.all.allocated", 0), st.get("segment.all.freed", 0),
                    st.get("reserved_bytes.all.current", 0) / 2 ** 30, st.get("active_bytes.all.current", 0) / 2 ** 30,
                    st.get("num_alloc_retries", 0)), file=sys.stderr, flush=True)
        e1 = marks[-1]
        host_ms.append((time.time() - w0) * 1e3 / k)        # host time to ENQUEUE one step (launch-bound if ~= the device time)
        barrier()
        w1 = time.time()
        # per-step device time (between the step-end events) and host enqueue time: shows whether a slow run is uniformly slow or has outliers
        per_step.append({"device_ms": [round(marks[i].elapsed_time(marks[i + 1]), 2) for i in range(k)],
                         "host_ms": [round(1e3 * (hmarks[i + 1] - hmarks[i]), 2) for i in range(k)]})
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), w0, w1

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # (under NCCL the first steps after start-up still grow NCCL's and the allocator's pools: at least 5 untimed steps there)
    n_warm = max(args.warmup, 5 if world > 1 else 3)
    for _ in range(n_warm - 1):
        step_device()
    # the step enqueues ~1 000 launches from Python; a generation-2 garbage collection over torch's heap inside the timed region
    # stalls the launch thread for tens of ms (seen as 64 vs 76 ms/step between otherwise identical runs): collect now and move
    # the survivors out of the collector's reach, as a training loop would after its first iterations
    gc.collect()
    gc.freeze()
    step_device()                                # last warm-up step, after the collection (the first step after it is the slow one)
    n0 = lib.pa2s_launch_count()
    seg0 = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
    ops.KernelTimers.reset(rank == 0)
    prof = os.environ.get("PA2S_PROFILE_RANGE") == "1"      # ncu --profile-from-start off: capture the timed steps only
    if prof:
        torch.cuda.profiler.start()
    ms, w0, w1 = timed(step_device, args.steps)
    if prof:
        torch.cuda.profiler.stop()
    launches = lib.pa2s_launch_count() - n0
    mallocs = torch.cuda.memory_stats(dev).get("segment.all.allocated", 0) - seg0
    ktimes = ops.KernelTimers.summary() if rank == 0 else {}
    ops.KernelTimers.reset(False)
    step_e2e()
    ms_e2e, _, w2 = timed(step_e2e, args.steps)
    ops.check_sync_flags()
    # host time to enqueue ONE step on an idle device (in the timed loop the host runs ahead until the launch queue is full and is
    # then paced by the device, so `host_enqueue_ms_per_step` ~ device time there): the launch-bound floor of the step
    barrier()
    t_h = time.time()
    step_device()
    host_unblocked = (time.time() - t_h) * 1e3
    barrier()
    clips = B * world * args.steps
    value = clips / (ms / 1e3)
    e2e = clips / (ms_e2e / 1e3)
    h2d = audio_h.numel() * 4 + sum(t.numel() * 8 for t in gt_h) + (n_h.numel() * 8 if finetune else 0)

    # secondary measurement in the default headline run: the same step with single-bf16 contractions (BASELINE configs[2])
    also = None
    if args.also_steps > 0 and args.precision == "bf16x3" and not finetune:
        ops.set_precision(train="bf16")
        for _ in range(3):
            step_device()
        ms_b, _, _ = timed(step_device, args.also_steps)
        ms_be, _, w2 = timed(step_e2e, args.also_steps)
        ops.set_precision(train=args.precision)
        cb = B * world * args.also_steps
        also = {"precision": "bf16", "dtype": DTYPE["bf16"], "steps": args.also_steps, "value": cb / (ms_b / 1e3), "unit": UNIT,
                "ms_per_step": ms_b / args.also_steps, "e2e_value": cb / (ms_be / 1e3), "note": PRECISION_NOTE["bf16"]}
    if rank == 0:
        sampler.stop()

    # data-parallel correctness: after the timed steps every rank must hold bit-identical parameters, optimizer state and BatchNorm
    # running statistics (they only ever see all-reduced gradients / statistics)
    rank_consistent = None
    if world > 1:
        bn = torch.cat([b.reshape(-1).float() for n_, b in model.named_buffers() if "running" in n_])
        sums = torch.stack([t.contiguous().view(torch.int32).to(torch.int64).sum() for t in (opt.flat, opt.square_avg, opt.acc_delta, bn)])
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        rank_consistent = all(bool((a == allsums[0]).all().item()) for a in allsums)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = {}
        try:        # per-launch dram__bytes_read+write from the committed `ncu --set full` capture (profiles/), keyed by timer name
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except Exception:
            pass
        roof, roof_all = roofline(ktimes, B, peaks, S, traffic)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cstep, threads, kind = reference_step_factory(args.cpu_clips, args.workload)
            t0 = time.time()
            cstep()
            dt = time.time() - t0
            cpu = {"value": args.cpu_clips / dt, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{args.cpu_clips} clip(s), one fwd+bwd+clip+Adadelta step of the same workload through {_kind_text(kind)} "
                             f"(torch CPU fp32 + float64 numpy VQT), {dt:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": DTYPE[args.precision],
            "data": "synthetic",
            "config": {"workload": workload_name(B, args.workload, args.precision), "precision": PRECISION_NOTE[args.precision],
                       "global_batch": B * world, "decoder_steps_per_forward": S,
                       "parallelism": f"dp{world}", "l2": "working set >> 126 MB L2 (4.4 GB of conv activations per step), no flush needed",
                       "kernel_ms": {k: round(v[1], 4) for k, v in sorted(ktimes.items())}},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "host_enqueue_ms_per_step": round(host_ms[0], 3), "host_enqueue_ms_idle_device": round(host_unblocked, 3), "per_step": {"device_resident": per_step[0], "e2e": per_step[1]},
            "gpu_launches": int(launches), "cuda_mallocs_in_timed_region": int(mallocs),
            "clocks": sampler.summary(w0, w2),
            "roofline": roof, "roofline_all": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for r in roof_all],
            "cpu_baseline": cpu}
        if also is not None:
            line["also"] = also
        if rank_consistent is not None:
            line["rank_consistent"] = rank_consistent
        print(json.dumps(line), flush=True)
    if world > 1:
        if not rank_consistent:
            raise RuntimeError("data-parallel ranks diverged: parameters / optimizer state / BatchNorm statistics differ across ranks")
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- B200 arm, greedy decode
def run_b200_infer(args):
    """BASELINE configs[4]: batched greedy hierarchical decode (pretrain.py:302-306 evaluate path) of synthetic 12-s clips,
    model.eval() => exact-fp32 kernels, 32 clips per GPU per step (8 steps = 256 clips), audio -> VQT -> ConvStack -> encoder ->
    5 bars x (398 + 189) note steps (a random-init model never emits <eos>) -> argmax/unpad token lists on the host."""
    import torch
    import torch.distributed as dist
    import models
    from piano_a2s_b200 import kern, ops
    from piano_a2s_b200.synthetic import make_audio
    from piano_a2s_b200.vqt import VQT
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:                                        # replicas only: the group exists for the barrier and the max over ranks
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    torch.manual_seed(1234)
    model = models.ScoreTranscription(**CFG).to(dev).eval()
    model.decoder.consume_python_rng = False             # the per-step coin count of models.py:404 is irrelevant without teacher forcing
    vqt = VQT().to(dev)
    audio_h = make_audio(B, N_SAMPLES, seed=1234 + rank).pin_memory()
    audio_d = audio_h.to(dev)
    d2h = [0]

    def step_device():
        with torch.no_grad():
            return model(vqt(audio_d).unsqueeze(1), device=dev)

    def step_e2e():
        with torch.no_grad():
            outs = model(vqt(audio_h.to(dev, non_blocking=True)).unsqueeze(1), device=dev)
            toks = kern.greedy_tokens(outs)              # argmax + first-<eos> on the device, token lists on the host
        d2h[0] = B * 5 * (8 * sum(CFG["max_length"]) + 2 * 4 + 2 * 8)      # int64 tokens + int32 lengths per staff, int64 key / time signature
        return toks

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), w0, time.time()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    from piano_a2s_b200._lib import lib
    n0 = lib.pa2s_launch_count()
    ops.KernelTimers.reset(rank == 0)
    ms, w0, _ = timed(step_device, args.steps)
    launches = lib.pa2s_launch_count() - n0
    ktimes = ops.KernelTimers.summary() if rank == 0 else {}
    ops.KernelTimers.reset(False)
    step_e2e()
    ms_e2e, _, w2 = timed(step_e2e, args.steps)
    if rank == 0:
        sampler.stop()
        ops.check_sync_flags()
        clips = B * world * args.steps
        S = 5 * (CFG["max_length"][0] + CFG["max_length"][1])
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roof, roof_all = roofline({k: v for k, v in ktimes.items() if k.startswith("note_decoder")}, B, peaks, S, n_dec=10)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cstep, threads, kind = reference_infer_factory(args.cpu_clips)
            t0 = time.time()
            cstep()
            dt = time.time() - t0
            cpu = {"value": args.cpu_clips / dt, "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{args.cpu_clips} clip(s), greedy decode of the same workload through {_kind_text(kind)} (torch CPU fp32 + "
                             f"float64 numpy VQT), {dt:.1f} s"}
        print(json.dumps({
            "metric": "clips_per_sec_greedy_decode", "value": clips / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"pretrain.yaml model, batched greedy hierarchical decode of synthetic 12-s clips, {B} clips/GPU per step "
                                   f"({B * world * args.steps} clips), eval mode, exact-fp32 kernels, audio -> VQT -> tokens",
                       "global_batch": B * world, "decoder_steps_per_forward": S, "parallelism": f"replicas{world}",
                       "l2": "inputs larger than L2 (5.9 GB of conv activations per step)",
                       "kernel_ms": {k: round(v[1], 4) for k, v in sorted(ktimes.items())}},
            "e2e": {"value": clips / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": audio_h.numel() * 4, "d2h_bytes_per_step": d2h[0],
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": sampler.summary(w0, w2),
            "roofline": roof, "roofline_all": [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()} for r in roof_all],
            "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}     # B200_PROFILING.md fallback


def algorithmic_work(B, S_total, n_dec_calls, n_dec_bwd_calls=None):
    """{timer name: (kernel, FLOP per launch, compulsory HBM bytes per launch)} at T=1201, F=480 -- SURVEY 8(d)'s per-clip
    figures x the B clips one launch processes.  Activations count as 4 B/element (fp32, or the two
    bf16 pieces the bf16x3 contractions read); FLOPs are the single-precision-equivalent 2*MAC count, NOT x3 for the split."""
    T, F = 1201, 480
    px = float(T * F * B)
    ch = [1, 20, 20, 40, 40]
    w = {}
    # VQT: algorithmic = direct form 4*T*sum(N_k) = 1.313 GFLOP/clip; the kernel contracts the dense 960 x 800 filter bank
    # (1.84 GFLOP/clip) once per bf16 piece product (x6 for the three-piece split): the tensor pipe sees 8.4x the algorithmic FLOPs
    w["vqt_filterbank"] = ("tc_gemm_tma_kernel VQT filterbank (+audio split), bf16x6, overlapping frame rows", 1.313e9 * B,
                           B * (0.77e6 + T * 960 * 4.0))
    w["conv1_fwd"] = ("conv1_fwd_kernel fp32 stream (conv1 + BatchNorm sums)", 2 * 9 * 1 * 20 * px, px * 4 * (1 + 20))
    w["conv1_wgrad"] = ("conv1_wgrad_kernel fp32 stream", 2 * 9 * 1 * 20 * px, px * 4 * (1 + 20 + 20))
    for i in (2, 3, 4):
        ci, co = ch[i - 1], ch[i]
        fl = 2.0 * 9 * ci * co * px
        w[f"conv{i}_planes"] = (f"planes_kernel<{ci}> relu(bn(y)) -> bf16 hi/lo planes", 2 * ci * px, px * ci * (4 + 4))
        w[f"conv{i}_fwd"] = (f"conv_tma3_kernel<{ci},{co}> tcgen05 implicit GEMM fwd, kx taps share the window read (+BN sums)", fl, px * 4 * (ci + co))
        w[f"conv{i}_dy_planes"] = (f"planes_bwd_kernel<{co}> BN/ReLU backward -> bf16 planes", 8 * co * px, px * co * (4 + 4 + 4))
        w[f"conv{i}_wgrad"] = (f"conv_wgrad_tma_kernel<{ci},{co}> tcgen05 weight gradient", fl, px * 4 * (ci + co))
        # (the data gradient also reads y of the layer below: its epilogue forms that layer's BatchNorm-backward sums)
        w[f"conv{i}_dgrad"] = (f"conv_tma3_kernel<{co},{ci}> tcgen05 data gradient (+BN-backward sums of the layer below)", fl, px * 4 * (ci + co + ci))
    M, K, N = float(B * T), 19200.0, 256.0
    lin = 2 * M * K * N
    # bf16x3 holds an fp32 operand as TWO bf16 pieces (hi, lo): 4 B/element read by the GEMM, 4 B read + 4 B written by the split
    w["out_linear_split"] = ("split_bf16_kernel (a4 = relu(bn4(y4)) -> bf16 hi/lo, W -> hi/lo)", 2 * M * K, (M * K + N * K) * (4 + 4))
    w["out_linear_fwd"] = ("tc_gemm_tma_kernel out Linear fwd", lin, (M * K + N * K) * 4 + M * N * 4)
    w["out_linear_wgrad"] = ("tc_gemm_tma_kernel out Linear weight gradient", lin, (M * K + M * N) * 4 + N * K * 4)
    w["out_linear_dgrad"] = ("tc_gemm_tma_kernel out Linear data gradient", lin, (M * N + N * K) * 4 + M * K * 4)
    # encoder BiGRU recurrence, one launch = one layer, both directions: 2 dirs x T steps x B x 2*768*256 FLOP; reads gi, writes out (+gates)
    w["encoder_gru_fwd"] = ("gru_seq_fwd_kernel (cluster-of-8 persistent BiGRU layer)", 2 * T * B * 2.0 * 768 * 256, B * T * 2 * (768 + 256 + 1024) * 4.0)
    w["encoder_gru_bwd"] = ("gru_seq_bwd2_kernel (row-owner reverse recurrence, st.async partial-sum exchange)", 2 * 2 * T * B * 2.0 * 768 * 256, B * T * 2 * (768 + 256 + 1024 + 768) * 4.0)
    # note decoder: per executed step and clip 6.27 MFLOP and 3.69 MB streamed (enc 2.46 MB + Ep 1.23 MB; L2-resident at B=16)
    steps = S_total / max(n_dec_calls, 1)
    steps_bwd = S_total / max(n_dec_bwd_calls or n_dec_calls, 1)
    # (`steps` = executed (bar, step) pairs per launch: a launch of the multi-sequence kernel decodes several bars of a staff)
    w["note_decoder_fwd"] = ("decm_fwd_kernel (all steps of the teacher-forced bars of one staff segment, cooperative)", 6.27e6 * B * steps, 3.69e6 * B * steps)
    w["note_decoder_bwd"] = ("decm_bwd_kernel (all bars of a staff; +dlogits, out-projection GEMM; the deferred dEp/dv kernel runs on an auxiliary stream, outside this timer)", 2 * 6.27e6 * B * steps_bwd, 3.69e6 * B * steps_bwd)
    return w


def roofline(ktimes, B, peaks, S_total, traffic=None, n_dec=None):
    """One entry per timed kernel group; the headline `roofline` object is the group with the largest share of the step.
    achieved = algorithmic FLOPs (or bytes) per launch / mean CUDA-event duration of the launch (events recorded on the
    launching stream inside the timed region); the bound reported is the roof that binds for the algorithmic numbers."""
    measured = bool(peaks.get("hbm_gbs"))
    pk = peaks if measured else FALLBACK_PEAKS
    hbm = float(pk.get("hbm_gbs") or FALLBACK_PEAKS["hbm_gbs"])
    tf = float(pk.get("bf16_tflops_sustained") or pk.get("bf16_tflops") or FALLBACK_PEAKS["bf16_tflops_sustained"])
    src = "MEASURED_PEAKS.json (hbm_gbs, bf16_tflops_sustained)" if measured else "fallback of B200_PROFILING.md (6.65 TB/s, 1.4 PFLOP/s sustained)"
    nsteps = max(ktimes.get("conv1_fwd", (1, 0))[0], 1)
    if n_dec is None:
        n_dec = ktimes.get("note_decoder_fwd", (10, 0))[0] / nsteps          # launches per step (average: it depends on the bar coins)
    n_dec_bwd = ktimes.get("note_decoder_bwd", (0, 0))[0] / nsteps or None
    work = algorithmic_work(B, S_total, n_dec, n_dec_bwd)
    rows = []
    for name, (n, ms) in ktimes.items():
        if name not in work or ms <= 0:
            continue
        kern, fl, by = work[name]
        t_tensor, t_hbm = fl / (tf * 1e12), by / (hbm * 1e9)
        bound = "tensor" if t_tensor > t_hbm else "hbm"
        if bound == "tensor":
            ach, peak, unit = fl / (ms * 1e-3) / 1e12, tf, "TFLOP/s"
        else:
            ach, peak, unit = by / (ms * 1e-3) / 1e9, hbm, "GB/s"
        rows.append({"timer": name, "kernel": kern, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                     "traffic": (traffic or {}).get(name), "launches_timed": n, "ms_per_launch": ms, "ms_total": n * ms,
                     "tflops_algorithmic": fl / (ms * 1e-3) / 1e12, "gbs_algorithmic": by / (ms * 1e-3) / 1e9})
    if not rows:
        return None, []
    rows.sort(key=lambda r: -r["ms_total"])
    top = dict(rows[0])
    top["peak_source"] = src
    return top, rows


if __name__ == "__main__":
    a = parse()
    if a.impl in ("reference", "reference-gpu"):
        run_reference(a)
    elif a.workload == "infer":
        run_b200_infer(a)
    else:
        run_b200(a)
