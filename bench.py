#!/usr/bin/env python
"""bench.py -- multimodal ATQ training throughput on B200 (BASELINE.json metric), with the roofline of the
dominant hot-path kernel and the REFERENCE'S OWN CPU path timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload flickr8k|vitb16] [--mode parity|fast]

N > 1 is launched by torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.
A "step" is one optimisation step of the named synthetic config (re-quantization of every ternary layer the
optimizer touched, forward, contrastive loss, backward, gradient all-reduce when N > 1, AdamW).

  value            whole-job samples/s with the batch already resident in HBM (BASELINE config 2, the headline)
  e2e              the same step with the batch copied from pinned host memory and the loss read back
  roofline         dominant own kernel inside the headline step (CUDA events around every C-ABI call)
  cpu_baseline     `--impl reference` run on the host cores: the UNMODIFIED reference (oracle/_ref: its own atq, models
                   and loss) on the same config; falls back to the CPU port only if the staged copy is missing
  dropin           the reference's own models bound to THIS repo's atq, eager, torch.optim.AdamW: nothing but the
                   reference's five public names is used -- the cost of staying strictly behind the boundary
  configs.vitb16   BASELINE config 4 (ViT-B/16-sized + 12-layer text tower, 512 samples per GPU) on the same build
  rooflines        own kernels at BASELINE config-3 (GEMM) / config-5 (quantize, pack) shapes, timed alone
  summary          the figures BASELINE's metric names, compact, LAST in the line
"""
from __future__ import annotations

import argparse
import contextlib
import dataclasses
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "atq-multimodal_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "multimodal ATQ train samples/sec"
UNIT = "samples/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
L2_NOTE = "flushed between timed steps (256 MB write, outside the per-step events)"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


def shared_config(cfg, n_gpus):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": cfg.name, "per_gpu_batch": cfg.batch, "global_batch": cfg.batch * n_gpus,
            "parallelism": f"dp{n_gpus}", "l2": L2_NOTE}


# ---------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------
class L2Flusher:
    def __init__(self, device, nbytes=256 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)

    def __call__(self):
        self.buf.zero_()  # 256 MB write > 126 MB L2


def to_device(batch, device):
    return tuple(t.to(device, non_blocking=True) for t in batch)


def channels_last_images(batch):
    img = batch[0]
    return (img.contiguous(memory_format=torch.channels_last) if img.dim() == 4 else img,) + tuple(batch[1:])


def max_over_ranks(x, device, world):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def timed_steps(step_fn, steps, flush, world):
    """K steps, each bracketed by CUDA events on the current stream; the L2 flush runs between steps,
    outside the events.  Returns total device milliseconds over the K steps (this rank)."""
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    barrier(world)
    for i in range(steps):
        flush()
        starts[i].record()
        step_fn(i)
        ends[i].record()
    barrier(world)
    return sum(s.elapsed_time(e) for s, e in zip(starts, ends))


# ---------------------------------------------------------------------------------------
# reference arm: the unmodified reference (oracle/_ref) on the host cores
# ---------------------------------------------------------------------------------------
def reference_throughput(cfg, steps, warmup):
    """-> (samples/s, s/step, kind, description).  kind "reference": the reference's own atq + models + loss + scheduler
    (staged copy, oracle/install_ref.py) driven exactly like train_multimodal.py; kind "port": the oracle restatement
    (only when the staged copy is missing, or for the ViT-sized config the reference has no model for)."""
    from oracle import ref_env
    from workloads import train as T
    torch.set_num_threads(os.cpu_count() or 1)
    batches = T.synthetic_batches(cfg, 2, seed=42)
    if cfg.image_tower == "resnet18" and ref_env.available():
        ref_env.activate("reference")
        from oracle import ref_tasks as R
        with contextlib.redirect_stdout(sys.stderr):  # the reference's constructors print
            model = R.build_retrieval(cfg.vocab, cfg.embed_dim, cfg.hidden_dim, seed=42)
            R.step_schedule(model, cfg.epoch, cfg.total_epochs, cfg.warmup_epochs)
        _, manager = R.build_loss(model, cfg.epoch, cfg.total_epochs)
        opt = R.make_optimizer(model, cfg.lr)
        model.train()
        step = lambda b: R.retrieval_step(model, manager, opt, b)  # noqa: E731
        kind = "reference"
        what = ("the unmodified reference (oracle/_ref: its atq, models.ATQMultimodalRetrieval, HardNegativeMiningInfoNCE, "
                "GradualQuantizationScheduler) stepped as train_multimodal.py:540-585 does")
    else:
        from oracle import policy as P
        from workloads import models as M
        model, _, manager = T.build_retrieval(M.oracle_layers(), cfg)
        P.scheduler_step(model, cfg.epoch, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs)
        opt = T.make_optimizer(model, cfg)
        model.train()
        step = lambda b: T.retrieval_step(model, manager, opt, b)  # noqa: E731
        kind = "port"
        what = "the CPU port of the reference algorithm (oracle/ layers in the harness model)"
    for i in range(warmup):
        float(step(batches[i % 2]).detach())
    t0 = time.perf_counter()
    for i in range(steps):
        float(step(batches[i % 2]).detach())
    dt = time.perf_counter() - t0
    return cfg.batch * steps / dt, dt / steps, kind, what


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, s_per_step, kind, what = reference_throughput(cfg, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(s_per_step * 1e3, 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(cfg, args.gpus),
            "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{args.steps} full optimisation steps (batch {cfg.batch}) of {what}; torch CPU, {cores} threads"},
            "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(args, workload):
    """The reference arm in a process of its own (it binds the name `atq` to the reference's package)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(args.cpu_steps), "--warmup", "1",
           "--workload", workload]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    env["CUDA_VISIBLE_DEVICES"] = ""
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
        line = json.loads(out.stdout.strip().splitlines()[-1])
        return line["cpu_baseline"]
    except Exception as exc:  # the baseline is a reported number, not a reason to lose the GPU line
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": f"failed: {exc!r}"[:200]}


# ---------------------------------------------------------------------------------------
# kernel-level rooflines measured live (CUDA events, this process)
# ---------------------------------------------------------------------------------------
_LEAD_CYCLES = 3_000_000  # ~1.5 ms at 1.965 GHz


def _device_spin(cycles):
    """Queue a one-thread spin of `cycles` SM clocks on the current stream (no memory traffic).  `torch.cuda._sleep` is a
    private torch API: without it the timings simply keep the host launch latency they had before."""
    spin = getattr(torch.cuda, "_sleep", None)
    if spin is not None:
        spin(int(cycles))


def _event_time(fn, iters, flush=None):
    """Median device time of fn() in ms.  Per iteration: L2 flush, then a device-side spin (`torch.cuda._sleep`, one
    thread, no memory traffic) that the host uses to enqueue fn() completely - tensor allocation, ctypes pointer tables and
    launch latency of a 60-layer batched call are ~0.3-0.5 ms of Python, more than several of the kernels measured here -
    so the two events bracket the kernels back to back, i.e. launch durations, not host latency."""
    fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush()
        _device_spin(_LEAD_CYCLES)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        times.append(s.elapsed_time(e))
    times.sort()
    return times[len(times) // 2]  # ms


def _event_time_rotating(fn_i, n_bufs, rounds, iters, flush=None):
    """Average launch duration (ms) of fn_i(i) over `rounds` back-to-back passes over `n_bufs` DISTINCT inputs whose total
    size exceeds L2 (so every launch streams its input from HBM), one event pair around all of them: the ~6 us that a
    single event-bracketed launch carries (launch latency, event resolution ~1 us) would otherwise be a third of a 64 MiB
    layer's 12-20 us.  Results are kept alive until the events are read, so outputs do not share (L2-resident) memory."""
    for i in range(n_bufs):
        fn_i(i)
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush()
        _device_spin(_LEAD_CYCLES)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        keep = [fn_i(i) for _ in range(rounds) for i in range(n_bufs)]
        e.record()
        torch.cuda.synchronize()
        del keep
        times.append(s.elapsed_time(e) / (rounds * n_bufs))
    times.sort()
    return times[len(times) // 2]


# nominal B200 peaks quoted by BASELINE.json's north_star (report fractions against both; `frac` = measured)
NOMINAL_HBM_GBS = 8000.0
NOMINAL_BF16_TFLOPS = 2250.0


TIMING_NOTE = ("median of the iterations; CUDA events on the launching stream; L2 flushed before each; host launch latency "
               "hidden behind a 1.5 ms device-side spin queued ahead of the start event")


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed `ncu --set full` captures (profiles/)
NCU_TRAFFIC = {}
try:
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as _f:
        NCU_TRAFFIC = json.load(_f)
except Exception:
    pass


def gemm_rooflines(device, peaks, flush):
    """BASELINE config 3: TernaryLinear / RPB at 4096^2 and 8192^2 weights, 8192 tokens; forward GEMM alone and the
    whole layer forward+backward (operand splits, bias / alpha reductions included; weights quantized once, H7)."""
    import atq
    import atq._engine as eng
    out = []
    tf = peaks["bf16_tflops"]
    src = "measured" if peaks["_source"] == "measured" else "fallback"
    prev = atq.get_gemm_mode()
    N = 8192
    for size in (4096, 8192):
        for kind, ratio in (("TernaryLinear", None), ("RPB0.2", 0.2)):
            torch.manual_seed(0)
            mod = (atq.TernaryLinear(size, size) if ratio is None else
                   atq.ResidualPrecisionBoostLinear(size, size, ratio, True, 0.3)).to(device)
            x = torch.randn(N, size, device=device)
            gy = torch.randn(N, size, device=device)
            xi = x.clone().requires_grad_(True)
            for mode in ("fast", "parity"):
                atq.set_gemm_mode(mode)
                mod._ops.key = None
                with torch.no_grad():
                    mod(x[:8])

                def fwd():
                    with torch.no_grad():
                        mod(x)

                def fwdbwd():
                    mod.zero_grad(set_to_none=True)
                    xi.grad = None
                    mod(xi).backward(gy)

                ms_f = _event_time(fwd, 5, flush)
                ms_fb = _event_time(fwdbwd, 5, flush)
                n_gemm = 2 if ratio is None else 3
                fl = 2.0 * N * size * size
                key = f"{kind} {size} {mode}"
                out.append({"kernel": "tgemm_kernel (+ operand split)", "workload": f"config3 {kind} {size}x{size}, {N} tokens, {mode}",
                            "bound": "tensor", "unit": "TFLOP/s", "peak": tf, "peak_kind": f"bf16 burst {src}",
                            "fwd_ms": round(ms_f, 4), "fwd_achieved": round(fl / ms_f / 1e9, 1), "fwd_frac": round(fl / ms_f / 1e9 / tf, 4),
                            "fwdbwd_ms": round(ms_fb, 4), "achieved": round(n_gemm * fl / ms_fb / 1e9, 1),
                            "frac": round(n_gemm * fl / ms_fb / 1e9 / tf, 4),
                            "frac_of_nominal_2250TF": round(n_gemm * fl / ms_fb / 1e9 / NOMINAL_BF16_TFLOPS, 4),
                            "fwd_frac_of_nominal_2250TF": round(fl / ms_f / 1e9 / NOMINAL_BF16_TFLOPS, 4), "traffic": NCU_TRAFFIC.get(key),
                            "useful_flops": f"{n_gemm} GEMMs x 2*N*K*M (hi/lo terms count 0)"})
            del mod, x, gy, xi
            torch.cuda.empty_cache()
    atq.set_gemm_mode(prev)
    return out


def attention_rooflines(device, peaks, flush):
    import atq
    from atq import attention as A
    out = []
    tf = peaks["bf16_tflops"]
    src = "measured" if peaks["_source"] == "measured" else "fallback"
    g = torch.Generator(device=device).manual_seed(0)
    b_, h_, l_ = 512, 12, 197
    q, k, v = (torch.randn(b_, l_, h_ * 64, device=device, generator=g).requires_grad_(True) for _ in range(3))
    dout = torch.randn(b_, l_, h_ * 64, device=device, generator=g)
    seed = torch.tensor([1], dtype=torch.int64, device=device)
    prev_mode = atq.get_gemm_mode()
    for mode in ("fast", "parity"):
        atq.set_gemm_mode(mode)
        with torch.no_grad():
            ms_f = _event_time(lambda: A.attention_core(q, k, v, h_, None, None, 0.1, True, seed=seed), 5, flush)
        o = A.attention_core(q, k, v, h_, None, None, 0.1, True, seed=seed)

        def bwd():
            q.grad = k.grad = v.grad = None
            o.backward(dout, retain_graph=True)
        ms_b = _event_time(bwd, 5, flush)
        # algorithmic bytes: forward reads q, k, v and writes out; backward reads q, k, v, out, dout and writes dq, dk, dv
        # (fp32 [B, L, E] tensors).  At L = 197, head_dim 64 the HBM time (1.24 GB / 2.48 GB) is above the tensor time of
        # the useful flops, so the second roofline ("hbm") is the binding one; both are reported.
        tensor_bytes = 4.0 * b_ * l_ * h_ * 64
        for name, ms, fl, nten, key in (("attention_fwd_kernel", ms_f, 4.0, 4, "attention fwd parity"),
                                        ("attention_bwd_kernel", ms_b, 10.0, 8, "attention bwd parity")):
            ach = fl * l_ * l_ * 64 * b_ * h_ / (ms * 1e-3) / 1e12
            gbs = nten * tensor_bytes / (ms * 1e-3) / 1e9
            out.append({"kernel": f"{name} {mode}", "workload": f"config4 attention core {b_}x{h_}x{l_}x64, dropout 0.1",
                        "bound": "tensor", "achieved": round(ach, 1), "peak": tf, "unit": "TFLOP/s", "frac": round(ach / tf, 4),
                        "peak_kind": f"bf16 burst {src}", "ms": round(ms, 4),
                        "hbm": {"achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 4),
                                "alg_bytes": nten * tensor_bytes},
                        "traffic": NCU_TRAFFIC.get(key) if mode == "parity" else None})
        del o
    atq.set_gemm_mode(prev_mode)
    return out


def streaming_rooflines(device, peaks, flush):
    """BASELINE config 5: one 4096x4096 layer, and the 1 B-weight layer list (60 layers, mixed-precision per-layer
    sparsities from MixedPrecisionATQ), thresholds batched."""
    import atq._engine as eng
    from atq.mixed_precision_atq import MixedPrecisionATQ
    out = []
    hbm = peaks["hbm_gbs"]
    src = "measured" if peaks["_source"] == "measured" else "fallback"

    def rec(kernel, workload, bpe, n, ms, traffic=None, note=None):
        ach = bpe * n / (ms * 1e-3) / 1e9
        r = {"kernel": kernel, "workload": workload, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s",
             "frac": round(ach / hbm, 4), "frac_of_nominal_8TBs": round(ach / NOMINAL_HBM_GBS, 4), "peak_kind": f"copy {src}",
             "ms": round(ms, 4), "alg_bytes_per_elem": bpe, "traffic": traffic}
        if note:
            r["note"] = note
        out.append(r)

    g = torch.Generator(device=device).manual_seed(0)
    M = K = 4096
    n = M * K
    NB, ROUNDS = 4, 2   # 4 distinct 64 MiB layers = 256 MiB > L2 (126 MB): every launch streams from HBM
    ws4 = [(torch.rand(M, K, device=device, generator=g) * 2 - 1) / K ** 0.5 for _ in range(NB)]
    thr4 = [eng.adaptive_threshold(w, 0.3) for w in ws4]
    wl = f"config5 one layer {M}x{K}"
    rot = f"average over {NB * ROUNDS} back-to-back launches on {NB} distinct layers ({NB * n * 4 >> 20} MiB > L2)"
    rec("select (exact k-th |W|)", wl, 4.0, n, _event_time_rotating(lambda i: eng.adaptive_threshold(ws4[i], 0.3), NB, ROUNDS, 5, flush),
        NCU_TRAFFIC.get("select 4096"), "4 B/elem over the whole select; " + rot)
    rec("ternarize_kernel -> 2-bit", wl, 4.25, n, _event_time_rotating(lambda i: eng.ternarize_pack2(ws4[i], thr4[i]), NB, ROUNDS, 5, flush),
        NCU_TRAFFIC.get("ternarize_pack 4096"), rot)
    packed4 = [eng.ternarize_pack2(w, t) for w, t in zip(ws4, thr4)]
    tern4 = [eng.unpack2(pk, n) for pk in packed4]
    rec("pack fp32 ternary -> 2-bit", wl, 4.25, n, _event_time_rotating(lambda i: eng.pack2_from_f32(tern4[i]), NB, ROUNDS, 5, flush), None, rot)
    del tern4
    rec("unpack2_kernel -> fp32", wl, 4.25, n, _event_time_rotating(lambda i: eng.unpack2(packed4[i], n), NB, ROUNDS, 5, flush),
        NCU_TRAFFIC.get("unpack 4096"), rot)
    # the same four kernels as ONE event-bracketed launch each (includes ~6 us of launch latency / event resolution)
    w, thr = ws4[0], thr4[0]
    single = {"select": _event_time(lambda: eng.adaptive_threshold(w, 0.3), 5, flush),
              "ternarize -> 2-bit": _event_time(lambda: eng.ternarize_pack2(w, thr), 5, flush),
              "unpack -> fp32": _event_time(lambda: eng.unpack2(packed4[0], n), 5, flush)}
    out.append({"kernel": "one event-bracketed launch each (for comparison)", "workload": wl,
                "us": {k: round(v * 1e3, 2) for k, v in single.items()},
                "frac": {k: round((4.0 if k == "select" else 4.25) * n / (v * 1e-3) / 1e9 / hbm, 4) for k, v in single.items()}})
    del ws4, thr4, packed4, w, thr
    # ---- 1 B weights
    names = ["image_encoder.layers.{i}.self_attn.q_proj", "text_encoder.layers.{i}.linear1", "text_projector.{i}",
             "encoder.ffn.intermediate.{i}", "image_encoder.layers.{i}.linear2", "text_encoder.attention_pool.{i}"]
    shapes = [(4096, 4096)] * 59 + [(2464, 4096)]
    ss = []
    for i in range(len(shapes)):
        nm = names[i % len(names)].format(i=i)
        _, s = MixedPrecisionATQ.calculate_quantization_params(None, nm, (0, 5, 9)[i % 3], 10, 0.3 if "image" in nm else 0.2)
        ss.append(s)
    ws = [(torch.rand(m, k, device=device, generator=g) * 2 - 1) / k ** 0.5 for m, k in shapes]
    total = sum(t.numel() for t in ws)
    wl = f"config5 {total} weights in {len(ws)} layers, per-layer sparsity from MixedPrecisionATQ"

    def quantize_pack():
        th = eng.adaptive_threshold_batched(ws, ss)
        return eng.ternarize_pack2_batched(ws, th)     # one launch for all layers

    packed = quantize_pack()
    rec("threshold, batched over layers", wl, 4.0, total, _event_time(lambda: eng.adaptive_threshold_batched(ws, ss), 5, flush))
    ms_qp = _event_time(quantize_pack, 5, flush)
    rec("quantize+pack (threshold + ternarize -> 2-bit)", wl, 8.25, total, ms_qp, None,
        "two-read bound 8.25 B/elem (one read for the order statistic, one for ternarize); single-read bound 4.25 -> "
        f"{round(4.25 * total / (ms_qp * 1e-3) / 1e9 / hbm, 4)} of peak")
    numels = [t.numel() for t in ws]
    th = eng.adaptive_threshold_batched(ws, ss)
    rec("ternarize -> 2-bit, batched over layers", wl, 4.25, total, _event_time(lambda: eng.ternarize_pack2_batched(ws, th), 5, flush))
    rec("unpack 2-bit -> fp32", wl, 4.25, total, _event_time(lambda: eng.unpack2_batched(packed, numels), 5, flush))
    tern, _ = eng.unpack2_batched(packed, numels)
    del ws
    rec("pack fp32 ternary -> 2-bit", wl, 4.25, total, _event_time(lambda: eng.pack2_from_f32_batched(tern), 5, flush))
    out.append({"kernel": "quantize+pack throughput", "workload": wl, "gelem_per_s": round(total / ms_qp / 1e6, 1)})
    del tern, packed
    torch.cuda.empty_cache()
    return out


class CallProfiler:
    """Brackets every C-ABI call with CUDA events (a separate, untimed pass of the same step)."""

    FLOPS = {"atq_tgemm": lambda a: 2.0 * a[0] * a[1] * a[2], "atq_tgemm_dw_masked": lambda a: 2.0 * a[0] * a[1] * a[2],
             "atq_tgemm_packed": lambda a: 2.0 * a[0] * a[1] * a[2], "atq_tgemm_absmax": lambda a: 2.0 * a[0] * a[1] * a[2]}

    def __init__(self):
        self.records = []

    def __enter__(self):
        import atq._native as nv
        self.nv, self.orig = nv, nv.call

        def wrapped(name, *args):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            self.orig(name, *args)
            e.record()
            self.records.append((name, tuple(a for a in args[1:4] if isinstance(a, int)), s, e))

        nv.call = wrapped
        return self

    def __exit__(self, *exc):
        self.nv.call = self.orig
        torch.cuda.synchronize()

    def summary(self):
        by = {}
        for name, key, s, e in self.records:
            d = by.setdefault(name, {"ms": 0.0, "calls": 0, "flops": 0.0})
            d["ms"] += s.elapsed_time(e)
            d["calls"] += 1
            if name in self.FLOPS and len(key) == 3:
                d["flops"] += self.FLOPS[name](key)
        return by


# ---------------------------------------------------------------------------------------
# one workload on this repo's path
# ---------------------------------------------------------------------------------------
def measure_workload(args, cfg, steps, warmup, ctx, use_graph, clock_sampler=None):
    """Builds the harness model of `cfg` on the B200 atq, runs W warm-up + K timed steps (resident batch), K e2e steps
    (pinned host batch in, loss out), then an untimed profiling pass.  Returns the result dictionary."""
    import atq
    import atq._native as nv
    from atq import parallel
    from atq.mixed_precision_atq import GradualQuantizationScheduler
    from atq.optim import FlatAdamW
    from workloads import train as T

    rank, world, device, peaks, flush = ctx["rank"], ctx["world"], ctx["device"], ctx["peaks"], ctx["flush"]
    model, _, manager = T.build_retrieval(atq, cfg)
    model.to(device).train()
    if cfg.image_tower == "resnet18":
        # cuDNN's tensor-core convolutions are NHWC: keep the fp32 trunk channels-last so no NCHW<->NHWC transposes
        # run around every convolution (caller-side layout, same maths)
        model.image_encoder.base_model.to(memory_format=torch.channels_last)
    GradualQuantizationScheduler(model, cfg.total_epochs, 0.3, 0.2, warmup_epochs=cfg.warmup_epochs).step(cfg.epoch)
    opt = T.make_optimizer(model, cfg, capturable=use_graph, fused=True, adamw_cls=None if args.torch_adamw else FlatAdamW)
    sync = None
    if world > 1:  # --sparse-grads: only the entries under each RPB precision_mask travel
        sync = parallel.FlatGradAllReduce(model.parameters(), sparse_masks=parallel.rpb_masks(model) if args.sparse_grads else None,
                                          overlap=not args.no_overlap_grads, bucket_bytes=args.bucket_mb << 20)
    gather = parallel.gather_embeddings if world > 1 else None

    pool = 4
    host = T.synthetic_batches(cfg, pool, seed=42 + rank, pin=False)
    if cfg.image_tower == "resnet18":
        host = [channels_last_images(b) for b in host]
    host = [tuple(t.pin_memory() for t in b) for b in host]
    resident = [to_device(b, device) for b in host]
    losses = []

    def eager_step(batch):
        return T.retrieval_step(model, manager, opt, batch, gather, sync, atq.prepare_quantization)

    if clock_sampler is not None:
        clock_sampler.start()  # samples from warm-up to the end of the e2e leg: the GPU is under load throughout
    run_step = eager_step
    for i in range(warmup):
        eager_step(resident[i % pool])
    torch.cuda.synchronize()
    gstep = None
    if use_graph:
        # the whole step (quantize + forward + loss + backward [+ all-reduce] + AdamW) as one CUDA graph
        gstep = T.GraphedRetrievalStep(model, manager, opt, resident[0], gather, sync, prepare=atq.prepare_quantization)
        run_step = gstep
        for i in range(2):
            run_step(resident[i % pool])
        torch.cuda.synchronize()

    def step_resident(i):
        losses.append(run_step(resident[i % pool]).detach())

    def step_e2e(i):
        if gstep is not None:
            loss = gstep(host[i % pool])          # pinned host -> static device buffers, then replay
        else:
            loss = eager_step(to_device(host[i % pool], device))
        losses.append(float(loss.detach()))       # device -> host read of the step's loss

    k0 = nv.kernel_launch_count()
    ms_total = timed_steps(step_resident, steps, flush, world)
    launches = nv.kernel_launch_count() - k0
    if gstep is not None:
        launches = gstep.own_kernels_per_replay * steps  # replays re-launch the captured kernels
    ms_total = max_over_ranks(ms_total, device, world)
    step_e2e(0)
    ms_e2e = max_over_ranks(timed_steps(step_e2e, steps, flush, world), device, world)
    sustained = None
    if args.sustained_steps > 0 and gstep is not None:
        ms_long = max_over_ranks(timed_steps(step_resident, args.sustained_steps, lambda: None, world), device, world)
        sustained = {"steps": args.sustained_steps, "value": round(cfg.batch * world * args.sustained_steps / (ms_long * 1e-3), 2),
                     "unit": UNIT, "timed_region_s": round(ms_long * 1e-3, 3), "l2": "not flushed (back-to-back replays)"}
    clocks = clock_sampler.stop() if clock_sampler is not None else None

    global_batch = cfg.batch * world
    res = {"value": round(global_batch * steps / (ms_total * 1e-3), 2), "unit": UNIT, "steps": steps, "warmup": warmup,
           "ms_per_step": round(ms_total / steps, 4),
           "e2e": {"value": round(global_batch * steps / (ms_e2e * 1e-3), 2), "unit": UNIT,
                   "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]), "d2h_bytes_per_step": 4,
                   "ms_per_step": round(ms_e2e / steps, 4)},
           "gpu_launches": int(launches), "clocks": clocks, "final_loss": round(float(losses[-1]), 5) if losses else None,
           "execution": (("whole step captured in one CUDA graph" + ("" if args.serial_towers or cfg.image_tower != "resnet18" else
                                                                      ", image / text towers on two streams (concurrent graph branches)"))
                         if use_graph else "eager"),
           "sustained": sustained}

    # ---- roofline of the dominant kernel of THIS library inside the step (separate untimed pass; every rank runs it
    # because the step contains collectives, rank 0 reports)
    if gstep is not None:
        gstep.release()
    prof = CallProfiler()
    with prof:
        for i in range(2):
            if use_graph:
                # a launch-bound step (that is why it is graphed): run eagerly, the GPU would idle between launches and each
                # event pair would bracket host launch latency, not kernel time.  A device-side spin queued ahead of the
                # pass lets the host enqueue the step first, so the calls execute back to back as they do in the graph.
                _device_spin(40e-3 * 1.9e9)
            eager_step(resident[i % pool])
    summ = prof.summary()
    own_ms = sum(d["ms"] for d in summ.values()) / 2
    tf_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    src = "measured" if peaks["_source"] == "measured" else "fallback"
    gemm_ms = sum(d["ms"] for n, d in summ.items() if d["flops"] > 0) / 2
    gemm_fl = sum(d["flops"] for n, d in summ.items() if d["flops"] > 0) / 2
    # dominant kernel of the path = the ternary GEMM (tgemm_kernel, reached through atq_tgemm / atq_tgemm_packed /
    # atq_tgemm_dw_masked): its calls are aggregated; a non-GEMM kernel is only named if no GEMM ran
    gemm_calls = sum(d["calls"] for n, d in summ.items() if d["flops"] > 0)
    if gemm_ms > 0:
        top = ("tgemm_kernel (atq_tgemm + atq_tgemm_absmax + atq_tgemm_dw_masked + atq_tgemm_packed)", {"ms": gemm_ms * 2, "flops": gemm_fl * 2, "calls": gemm_calls})
    else:
        top = max(summ.items(), key=lambda kv: kv[1]["ms"]) if summ else None
    if top is not None:
        name, d = top
        if d["flops"] > 0:
            ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
            res["roofline"] = {"kernel": name, "bound": "tensor", "achieved": round(ach, 3), "peak": tf_peak, "unit": "TFLOP/s",
                               "frac": round(ach / tf_peak, 5), "traffic": None, "peak_kind": f"bf16 sustained {src}",
                               "calls_per_step": d["calls"] // 2, "ms_per_step": round(d["ms"] / 2, 4),
                               "share_of_own_kernel_time": round(d["ms"] / 2 / max(own_ms, 1e-9), 3),
                               "note": "useful flops 2*rows*cols*k per call, CUDA events around each call of an eager pass" +
                                       (" queued behind a 40 ms device-side spin (host launch latency hidden)" if use_graph else "")}
        else:
            res["roofline"] = {"kernel": name, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": None, "traffic": None, "ms_per_step": round(d["ms"] / 2, 4)}
    if gemm_ms > 0:
        res["gemm_in_step"] = {"useful_tflops": round(gemm_fl / (gemm_ms * 1e-3) / 1e12, 2),
                               "frac_of_sustained_peak": round(gemm_fl / (gemm_ms * 1e-3) / 1e12 / tf_peak, 4),
                               "ms_per_step": round(gemm_ms, 3), "share_of_step": round(gemm_ms / (ms_total / steps), 3)}
    res["own_kernel_ms_per_step"] = round(own_ms, 4)
    res["own_calls"] = {k: {"calls_per_step": v["calls"] // 2, "ms_per_step": round(v["ms"] / 2, 4)} for k, v in
                        sorted(summ.items(), key=lambda kv: -kv[1]["ms"])[:8]}
    # release everything this workload holds (the next one needs the memory)
    if gstep is not None:
        gstep.graph.reset()
    del gstep, model, opt, sync, resident, host, manager
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def measure_dropin(args, cfg, steps, warmup, ctx):
    """The reference's OWN model / loss / scheduler / step (oracle/_ref models + utils) bound to this repo's atq:
    eager, torch.optim.AdamW, no prepare_quantization, no fused FFN / attention / residual, no CUDA graph."""
    from oracle import ref_env
    if not ref_env.available():
        return {"unavailable": "oracle/_ref not staged"}
    ref_env.activate("b200")
    from oracle import ref_tasks as R
    from workloads import train as T
    device, flush = ctx["device"], ctx["flush"]
    with contextlib.redirect_stdout(sys.stderr):
        model = R.build_retrieval(cfg.vocab, cfg.embed_dim, cfg.hidden_dim, seed=42)
        R.step_schedule(model, cfg.epoch, cfg.total_epochs, cfg.warmup_epochs)
    model.to(device).train()
    _, manager = R.build_loss(model, cfg.epoch, cfg.total_epochs)
    opt = R.make_optimizer(model, cfg.lr)
    host = [tuple(t.pin_memory() for t in b) for b in T.synthetic_batches(cfg, 4, seed=42)]
    resident = [to_device(b, device) for b in host]
    for i in range(warmup):
        R.retrieval_step(model, manager, opt, resident[i % 4])
    ms = timed_steps(lambda i: R.retrieval_step(model, manager, opt, resident[i % 4]), steps, flush, 1)
    ms_e2e = timed_steps(lambda i: float(R.retrieval_step(model, manager, opt, to_device(host[i % 4], device)).detach()), steps, flush, 1)
    out = {"value": round(cfg.batch * steps / (ms * 1e-3), 2), "unit": UNIT, "ms_per_step": round(ms / steps, 4),
           "e2e": round(cfg.batch * steps / (ms_e2e * 1e-3), 2),
           "what": "reference models.ATQMultimodalRetrieval + HardNegativeMiningInfoNCE + train_multimodal.py step, unmodified, on this atq (eager)"}
    del model, opt, resident, host
    torch.cuda.empty_cache()
    return out


def run_ours(args, cfg):
    import atq
    from atq import parallel
    from workloads import train as T

    rank, world, local = parallel.init_from_env()
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device; there is no CPU fallback"
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    atq.set_gemm_mode(args.mode)
    if args.serial_towers:
        from workloads import models as WM
        WM.PARALLEL_TOWERS = False
    if args.torch_loss:
        T.FUSED_LOSS = False
    ctx = {"rank": rank, "world": world, "device": device, "peaks": load_peaks(), "flush": L2Flusher(device)}
    peaks = ctx["peaks"]

    # the small-shape config is launch-bound and runs as one CUDA graph; the ViT-B-sized config is kernel-bound (and
    # its activations would be held twice by a capture pool), so it runs eagerly
    use_graph = (not args.no_graph) and cfg.image_tower == "resnet18"
    head = measure_workload(args, cfg, args.steps, args.warmup, ctx, use_graph, ClockSampler(local))

    line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": shared_config(cfg, world),
            "arm": {"gemm_mode": args.mode,
                    "gemm_arithmetic": {"parity": "scaled fp16 hi+lo operand pairs (per-tensor power-of-two scale), two fp32 TMEM accumulators per tile",
                                        "parity_bf16": "bf16 hi+lo operand pairs, two fp32 TMEM accumulators per tile",
                                        "fast": "bf16 operands, fp32 TMEM accumulate"}[args.mode],
                    "quant_schedule": f"GradualQuantizationScheduler epoch {cfg.epoch}/{cfg.total_epochs}", "execution": head["execution"]},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "final_loss": head["final_loss"]}
    if "roofline" in head:
        line["roofline"] = head["roofline"]
    for k in ("gemm_in_step", "own_kernel_ms_per_step", "own_calls", "sustained"):
        if head.get(k) is not None:
            line[k] = head[k]

    extra = {}
    if args.workload == "flickr8k" and not args.no_vitb16:
        # BASELINE config 4 on the same build, same process, same ranks (eager; fewer steps: ~0.3 s each)
        vsteps = min(args.steps, args.vit_steps)
        v = measure_workload(args, T.VITB16, vsteps, 3, ctx, False, None)
        extra["vitb16"] = {"workload": T.VITB16.name, "value": v["value"], "unit": UNIT, "ms_per_step": v["ms_per_step"],
                           "steps": vsteps, "e2e": v["e2e"]["value"], "gpu_launches": v["gpu_launches"],
                           "roofline": v.get("roofline"), "gemm_in_step": v.get("gemm_in_step"), "own_calls": v["own_calls"],
                           "final_loss": v["final_loss"], "mode": args.mode}
    if rank == 0:
        if world == 1 and not args.no_kernel_rooflines:
            flush = ctx["flush"]
            rl = attention_rooflines(device, peaks, flush) + streaming_rooflines(device, peaks, flush) + gemm_rooflines(device, peaks, flush)
            line["rooflines"] = rl
            line["rooflines_timing"] = TIMING_NOTE
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_subprocess(args, args.workload)
        if world == 1 and not args.no_dropin and cfg.image_tower == "resnet18":
            try:
                line["dropin"] = measure_dropin(args, cfg, args.steps, args.warmup, ctx)
            except Exception as exc:
                line["dropin"] = {"unavailable": repr(exc)[:300]}
        if extra:
            line["configs"] = extra
        # ---- compact summary, last in the line
        summary = {"config2_samples_per_s": line["value"], "config2_e2e": line["e2e"]["value"]}
        if "vitb16" in extra:
            summary["config4_samples_per_s"] = extra["vitb16"]["value"]
            summary["config4_ms_per_step"] = extra["vitb16"]["ms_per_step"]
            if extra["vitb16"].get("gemm_in_step"):
                summary["config4_gemm_frac_in_step"] = extra["vitb16"]["gemm_in_step"]["frac_of_sustained_peak"]
        if "dropin" in line and "value" in line["dropin"]:
            summary["dropin_samples_per_s"] = line["dropin"]["value"]
        if line.get("cpu_baseline", {}).get("value"):
            summary["reference_cpu_samples_per_s"] = line["cpu_baseline"]["value"]
            summary["reference_kind"] = line["cpu_baseline"]["kind"]
        for r in line.get("rooflines", []):
            if r.get("bound") == "tensor" and "fwdbwd_ms" in r:
                summary["c3 " + r["workload"].split("config3 ")[1].replace(", 8192 tokens", "") + " fwd/fwdbwd frac"] = [r["fwd_frac"], r["frac"]]
            elif r.get("bound") == "hbm":
                summary["c5 " + r["kernel"] + (" 1B" if "layers" in r["workload"] else " 4096^2")] = r["frac"]
        line["summary"] = summary
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear down in a fixed order; a watchdog guarantees the process exits even if communicator destruction stalls
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        torch.cuda.synchronize()
        torch.distributed.barrier()
        torch.cuda.synchronize()
        try:
            torch.distributed.destroy_process_group()
        finally:
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["flickr8k", "vitb16"], default="flickr8k")
    ap.add_argument("--mode", choices=["parity", "parity_bf16", "fast"], default="parity")
    ap.add_argument("--batch", type=int, default=None, help="override per-GPU batch (debug only; invalidates the number)")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--vit-steps", type=int, default=8, help="timed steps of the config-4 sub-run")
    ap.add_argument("--sustained-steps", type=int, default=400, help="extra back-to-back graph replays (timed region >= 1 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of as one CUDA graph")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-vitb16", action="store_true", help="skip the config-4 sub-run")
    ap.add_argument("--no-dropin", action="store_true", help="skip the reference-models-on-this-atq arm")
    ap.add_argument("--serial-towers", action="store_true", help="run the image and text towers on one stream (no graph branch concurrency)")
    ap.add_argument("--torch-loss", action="store_true", help="contrastive loss as torch ops instead of the fused kernels")
    ap.add_argument("--torch-adamw", action="store_true", help="use torch.optim.AdamW(fused=True) instead of atq.optim.FlatAdamW")
    ap.add_argument("--no-overlap-grads", action="store_true", help="N>1: one all-reduce after backward instead of overlapped buckets")
    ap.add_argument("--bucket-mb", type=int, default=24, help="N>1: gradient bucket size of the overlapped all-reduce")
    ap.add_argument("--sparse-grads", action="store_true", help="N>1: all-reduce only the masked entries of RPB weight gradients")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    from workloads import train as T
    cfg = T.FLICKR8K_SHAPE if args.workload == "flickr8k" else T.VITB16
    if args.batch:
        cfg = dataclasses.replace(cfg, batch=args.batch, name=cfg.name + f" [DEBUG batch {args.batch}]")
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
