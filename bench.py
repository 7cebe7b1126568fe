#!/usr/bin/env python
"""Benchmark of the batched-evaluation hot path (BASELINE.json metric: eval images/s at 224x224).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

A "step" is one batch through the whole path: fused resize/normalise of ``batch`` ISIC-shaped
600x450 uint8 images -> SkinCancerListModel forward (bf16 tensor cores, fp32 accumulate) -> per-group
confusion counts.  N=1 workload = BASELINE configs[1] (batch 256, bf16, 1 x B200).  For N>1 the driver
launches one rank per GPU with torchrun; every rank evaluates its own batches (weak scaling) and the
only collective is ONE all-reduce of the 576-byte count tensor at the end of the timed region.

Prints one JSON line (rank 0).  ``value`` = whole-job images/s with inputs resident in HBM,
CUDA-event timed, max over ranks.  ``e2e`` = same metric through the public ``EvalEngine.step`` call
from pinned HOST buffers (H2D of every batch and D2H of the counts inside the timed region).
``--impl reference`` times the reference's own CPU path (the oracle port of it: scipy/numpy resize +
fp32 torch model + dict-based analysis) on this box's host cores with the same JSON schema and the same
``config`` (256 images per step, resize workers overlapped with the forward pass like the reference's DataLoader).
``--workload shard1m`` is BASELINE configs[2]: 1 M logical images sharded over the ranks by whole batches, ONE
all-reduce, and rank 0 re-counts every image alone afterwards and asserts the reduced counts are bit-identical.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SRC_H, SRC_W, OUT = 450, 600, 224
METRIC, UNIT = "eval_images_per_sec_224", "images/s"
NOMINAL_BF16_TFLOPS = 2250.0          # dense bf16, B200 (the profiling guide); only used when a stage beats the measured GEMM
WORKLOAD = ("configs[1]: SkinCancerListModel eval, batch 256 per GPU, bf16, fused resize+normalise from synthetic "
            "600x450 uint8 ISIC-shaped images, per-Fitzpatrick-group confusion counts")
# --workload: the default is the configuration the metric is quoted on; the others are BASELINE configs[3] / [4]
# (alternate architecture, high-resolution input) measured with the same harness.
WORKLOADS = {
    "list224": dict(kind="SkinCancerListModel", out=224, batch=256, text=WORKLOAD),
    "fourconv224": dict(kind="SkinCancerModel", out=224, batch=512,
                        text="configs[3]: SkinCancerModel (= jgi_hiba_2022_model) eval, batch 512 per GPU, bf16, fused "
                             "resize+normalise from synthetic 600x450 uint8 images, per-group confusion counts"),
    "optuna224": dict(kind="optuna_best", out=224, batch=128,
                      text="tone_bias_optuna.create_best_model (conv 192/172/22/86, linear 227/80/86; 10.5 GFLOP per "
                           "image) eval, batch 128 per GPU, bf16, zero-padded channel buffers"),
    "shard1m": dict(kind="SkinCancerListModel", out=224, batch=256, images=1_000_000,
                    text="configs[2]: 1M synthetic logical images (600x450 uint8 pixels from a ring of 4 distinct "
                         "batches, per-index metadata) sharded across the GPUs by whole batches of 256, "
                         "SkinCancerListModel bf16, one NCCL all-reduce of the per-group counts, reduced counts "
                         "compared bit for bit with a 1-GPU recount of all images"),
    "list512": dict(kind="SkinCancerListModel", out=512, batch=128,
                    text="configs[4]: SkinCancerListModel eval at 512x512 (first Linear 524288->512), batch 128 per "
                         "GPU, bf16, fused resize+normalise from synthetic 600x450 uint8 images"),
}


def conv_flops_per_image(cin: int, cout: int, k: int, side: int) -> int:
    return 2 * side * side * cout * cin * k * k


def pre_bytes_per_image(out: int) -> int:
    return SRC_H * SRC_W * 3 + 3 * out * out * 2                    # algorithmic: u8 in + 3-channel bf16 out


def measured_peaks():
    """HBM copy GB/s and bf16 GEMM TFLOP/s of this pool's B200s (driver-written file; else the profiling guide's
    fallback).  `bf16_tflops` is the burst figure (a kernel timed alone -- what the per-stage breakdown does),
    `bf16_tflops_sustained` the back-to-back figure."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(stage: str, batch: int):
    """DRAM bytes per launch of `stage` (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    `ncu --set full` capture of the same workload (profiles/r02_traffic.json if present, else the round-1 file; per
    image there), scaled to this run's batch; None when no capture is committed."""
    for name in ("r02_traffic.json", "r01_final_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                per_image = json.load(f)["dram_bytes_per_image"]
            return per_image[stage] * batch
        except Exception:
            continue
    return None


class ClockSampler:
    """SM clock and throttle reasons WHILE the timed region runs, without touching it: the host issues the K steps of
    the region far ahead of the GPU (a step is launched in ~20 us and runs for ~450 us), so rank 0 polls NVML from the
    MAIN thread only after every launch of the region has been issued, until the closing event has completed
    (``sample_until``).  A polling thread -- or eight ranks polling -- delays kernel launches: NVML queries take the
    driver for milliseconds on an 8-GPU box and cost 35 % of the timed region there.  Without NVML the fallback is
    ``nvidia-smi -lms 100`` in a subprocess."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.nvml = None
        self.enabled = enabled
        if not enabled:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                pr = torch.cuda.get_device_properties(index)
                bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        if self.nvml is None:                      # fallback: start nvidia-smi here too, i.e. before the barrier
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read, daemon=True)
                self.thread.start()
            except Exception:
                self.proc = None

    def _poll_once(self):
        n = self.nvml
        names = (n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
                 n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap)
        try:
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            self.rows.append([mhz, self.max_mhz] + ["Active" if mask & bit else "Not Active" for bit in names])
        except Exception:
            pass

    def sample_until(self, event, max_seconds: float = 120.0):
        """Call after the last launch of the timed region: polls until ``event`` (recorded at its end) has completed."""
        if not self.enabled or self.nvml is None:
            return
        t0 = time.perf_counter()
        while True:
            self._poll_once()                      # at least one sample, taken while the GPU still works on the region
            if event.query() or time.perf_counter() - t0 > max_seconds:
                break
            time.sleep(0.002)

    def __enter__(self):
        return self

    def _read(self):
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 7:
                self.rows.append([c[0], c[1]] + c[3:7])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if str(v).lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------------
def _resize_one(im):
    from oracle import resize as R
    return R.transform_u8(im, (OUT, OUT), R.resize_scipy)


def workload_config(wl: dict, batch: int, world: int, n_slots: int = 4) -> dict:
    """The ``config`` object of the JSON line -- the same for both arms (the reference arm runs 'your arm's config')."""
    return {"workload": wl["text"], "global_batch": batch * world, "batch_per_gpu": batch, "parallelism": f"dp{world}",
            "l2": f"inputs larger than L2: ring of {n_slots} distinct {batch * SRC_H * SRC_W * 3 / 1e6:.0f} MB batches "
                  "per GPU"}


def _cpu_inputs(n_images: int, seed: int):
    from skin_image_analysis_b200.synthetic import counter_metadata
    rng = np.random.default_rng(seed)
    imgs = rng.integers(0, 256, (n_images, SRC_H, SRC_W, 3), dtype=np.uint8)
    return imgs, counter_metadata(np.arange(n_images) + seed * n_images, 1)


def _cpu_forward_and_analysis(x, meta, state):
    from oracle import analysis as oa
    from oracle import model as om
    label, ftype, sex, control = meta
    logp = om.forward(om.LIST_MODEL, state, x)
    pred = om.predict(logp).numpy()
    names, fitz = ["benign", "malignant"], ["I", "II", "III", "IV", "V", "VI"]
    inst = {i: {"benign_malignant": names[label[i]], "prediction": names[pred[i]], "skin_type": fitz[ftype[i]],
                "skin_tone": "light" if ftype[i] < 2 else "dark", "sex": ["male", "female"][sex[i]],
                "control": ["rich", "poor"][control[i]], "age": 50.0} for i in range(len(pred))}
    try:
        oa.analyse_predictions(inst, out=lambda *a: None)
    except ZeroDivisionError:
        pass                                       # a tiny sample may leave a group empty, as in the reference


def cpu_reference_steps(n_steps: int, n_images: int, seed0: int, state, pool):
    """n_steps of the reference's path for n_images each on the host, pipelined the way the reference runs it: the
    transform of batch k+1 happens in worker PROCESSES (DataLoader workers, tone_bias_test.py:637; scipy.ndimage holds
    the GIL so threads do not scale) while the main process runs the fp32 torch forward + the dict-based analysis of
    batch k on all cores.  Returns the wall time of the n_steps (the first transform is inside it)."""
    inputs = [_cpu_inputs(n_images, seed0 + s) for s in range(n_steps)]          # synthetic data: outside the clock
    t0 = time.perf_counter()
    pending = pool.map_async(_resize_one, list(inputs[0][0]))
    for s in range(n_steps):
        x = torch.from_numpy(np.stack(pending.get()))
        if s + 1 < n_steps:
            pending = pool.map_async(_resize_one, list(inputs[s + 1][0]))
        _cpu_forward_and_analysis(x, inputs[s][1], state)
    return time.perf_counter() - t0


def cpu_baseline():
    """Runs the reference arm in a fresh process (no CUDA context there, so worker processes can fork) on a bounded
    sample: 3 steps of 256 images."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        return line["cpu_baseline"]
    except Exception as exc:                                        # never let the baseline kill the bench line
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = min(cores, 32)
    pool = mp.get_context("fork").Pool(workers)         # before any torch op starts its thread pool
    pool.map(_resize_one, [np.zeros((SRC_H, SRC_W, 3), np.uint8)] * workers)
    from oracle import model as om
    torch.set_num_threads(cores)
    state = om.synthetic_state_dict(om.LIST_MODEL, seed=0)
    wl = WORKLOADS["list224"]
    sample = args.ref_sample or wl["batch"]
    cpu_reference_steps(max(1, args.warmup), sample, 100, state, pool)
    t = cpu_reference_steps(args.steps, sample, 200, state, pool)
    pool.close()
    value = sample * args.steps / t
    desc = (f"{sample} synthetic 600x450 images per step x {args.steps} steps: scipy.ndimage resize in {workers} "
            f"worker processes overlapped with the fp32 torch CPU forward on {cores} threads + dict-based analysis; "
            "the reference is pure Python with un-installable imports (skimage, optuna) and /root/reference does not "
            "exist on the GPU box, so this is the oracle PORT of it, not the reference's own files")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, wl["batch"], max(world, args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def build_state(device, centre: bool = True, kind: str = "SkinCancerListModel", out: int = OUT):
    """Random-init weights of the reference architecture (xavier-normal weights, default biases) with the
    head centred so both classes occur (un-centred random init predicts one class)."""
    from skin_image_analysis_b200 import ops
    from skin_image_analysis_b200.engine import plan_from_state_dict
    from skin_image_analysis_b200.synthetic import random_state_dict
    state = random_state_dict(kind, out, seed=0)
    if not centre:
        return state
    plan = plan_from_state_dict(state, device)
    # calibrate on inputs from the same distribution as the benchmark's (resized uint8 noise)
    g = torch.Generator(device=device).manual_seed(1)
    u8 = torch.randint(0, 256, (64, SRC_H, SRC_W, 3), dtype=torch.uint8, device=device, generator=g)
    logp, _pred = plan.forward_nhwc4(ops.preprocess_u8hwc(u8, (out, out), ops.LAYOUT_NHWC4_BF16))
    last = [k for k in state if k.endswith(".bias")][-1]
    state[last][1] -= float((logp[:, 1] - logp[:, 0]).median())
    del plan
    return state


def tensor_peak(peaks, clocks):
    """Which measured bf16 GEMM figure a tensor-bound kernel of this run is held against.  MEASURED_PEAKS.json has
    two: ``bf16_tflops`` (burst: best of 10 at the maximum SM clock) and ``bf16_tflops_sustained`` (4 s back to back,
    power-capped clocks).  A region that ran at the maximum clock with no sw_power_cap gets the BURST figure; only a
    region the clock record shows capped (or clocked clearly below max) gets the sustained one."""
    sm, mx = clocks.get("sm_mhz"), clocks.get("sm_max_mhz")
    capped = "sw_power_cap" in clocks.get("reasons", []) or (sm is not None and mx and sm < 0.97 * mx)
    if capped:
        return peaks["bf16_tflops_sustained"], "sustained bf16 GEMM (clock record shows a power cap / reduced clock)"
    return peaks["bf16_tflops"], "burst bf16 GEMM (region ran at the maximum SM clock, no power cap)"


def stage_breakdown(eng, batch, peaks, tensor_tflops, iters=12, done_event=None, sampler=None):
    """Average duration of every kernel of the step, measured INSIDE whole steps: the kernels are launched
    one after the other on the engine stream (no graph) with a CUDA event between consecutive launches,
    input slots rotating as in the timed region, all iterations queued ahead of the GPU.  Roofline fractions use the measured peaks: HBM copy for the
    preprocess kernel and for fc1 (a 103 MB weight stream per launch), ``tensor_tflops`` for the conv kernels."""
    from skin_image_analysis_b200 import ops
    plan, ws = eng.plan, eng.plan.workspace(batch)
    acts = ws["acts"]
    size = eng.out_size
    names = ["preprocess"] + [f"conv{i + 1}" for i in range(len(plan.convs))] + ["fc1", "tail"]

    def launch_all(slot, ev):
        k = 0
        ev[k].record(eng.stream); k += 1
        ops.preprocess_u8hwc(eng.u8[slot], (size, size), ops.LAYOUT_NHWC4_BF16, out=eng.x4)
        ev[k].record(eng.stream); k += 1
        h = eng.x4
        for i in range(len(plan.convs)):
            h = plan.conv_block(i, h, acts[i])
            if plan.needs_pad[i]:
                v = plan.valid[i + 1]
                h = ops.pad_nhwc(h, (v, v), (plan.in_hw[i + 1],) * 2, out=ws["padded"][i])
            ev[k].record(eng.stream); k += 1
        ops.linear_splitk(h.view(batch, -1), plan.w1, ws["splits"], out=ws["partial"])
        ev[k].record(eng.stream); k += 1
        plan.tail(ws["partial"], logp=ws["logp"], pred=ws["pred"])
        ev[k].record(eng.stream)

    # All iterations are queued back to back and read after ONE synchronize: with a synchronize per iteration the GPU
    # idles before every first launch, and the first stage (preprocess) absorbs the host's launch latency (~17 us).
    total = dict.fromkeys(names, 0.0)
    warm = 3
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)] for _ in range(iters + warm)]
    with torch.cuda.stream(eng.stream):
        for it in range(iters + warm):
            launch_all(it % eng.n_slots, evs[it])
        if done_event is not None and sampler is not None:
            done_event.record(eng.stream)
            sampler.sample_until(done_event)         # clocks while the burst runs
        eng.stream.synchronize()
    for it in range(warm, iters + warm):
        for j, nm in enumerate(names):
            total[nm] += evs[it][j].elapsed_time(evs[it][j + 1]) * 1e-3 / iters
    out = {}
    t = total["preprocess"]
    gbs = pre_bytes_per_image(size) * batch / t / 1e9
    out["preprocess"] = {"ms": t * 1e3, "bound": "hbm", "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                         "frac": gbs / peaks["hbm_gbs"], "bytes_per_image": pre_bytes_per_image(size)}
    side = size
    flops = {}
    cin = 3
    for i, cout in enumerate(plan.widths):                    # algorithmic FLOPs: real widths, no padding
        flops[f"conv{i + 1}"] = conv_flops_per_image(cin, cout, 7 if i == 0 else 3, side)
        side //= 2
        cin = cout
    for nm, fl in flops.items():
        tf = fl * batch / total[nm] / 1e12
        out[nm] = {"ms": total[nm] * 1e3, "bound": "tensor", "achieved": tf, "unit": "TFLOP/s",
                   "peak": tensor_tflops, "frac": (tf / tensor_tflops) if tensor_tflops else None, "flops_per_image": fl}
    # fc1: the weight matrix (bf16, read once) + the activations + the fp32 split-K partials -- HBM-bound at this batch
    feat = plan.widths[-1] * side * side
    fc1_bytes = plan.n1 * feat * 2 + batch * feat * 2 + ws["splits"] * batch * plan.n1_pad * 4
    gbs = fc1_bytes / total["fc1"] / 1e9
    out["fc1"] = {"ms": total["fc1"] * 1e3, "bound": "hbm", "achieved": gbs, "unit": "GB/s", "peak": peaks["hbm_gbs"],
                  "frac": gbs / peaks["hbm_gbs"], "bytes_per_launch": fc1_bytes,
                  "tflops": 2 * feat * plan.n1 * batch / total["fc1"] / 1e12}
    out["tail"] = {"ms": total["tail"] * 1e3}
    return out


def pinned_h2d_peak(dev, nbytes: int, reps: int = 3, chain: int = 6):
    """GB/s of plain pinned-host -> device copies of nbytes on this GPU (best of reps; each rep = `chain` copies back
    to back between two events, so the start-up latency of a single copy does not understate the link): what the e2e
    figure is held against."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    best = 0.0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in range(chain):
            dst[c % 2].copy_(src, non_blocking=True)
        e1.record()
        e1.synchronize()
        best = max(best, chain * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    del src, dst
    return best


def run_ours(args):
    from skin_image_analysis_b200 import distributed as D
    from skin_image_analysis_b200 import tone_bias_test as tt
    from skin_image_analysis_b200.engine import EvalEngine, N_ATTR
    from skin_image_analysis_b200.synthetic import counter_metadata, device_u8_batches
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = D.bind_to_gpu_numa_node(local) if world > 1 else {"numa_node": None, "cpus": 0}
    wl = WORKLOADS[args.workload]
    batch, warmup = (args.batch or wl["batch"]), args.warmup
    sharded = "images" in wl                       # configs[2]: a fixed number of logical images, strong scaling
    peaks = measured_peaks()

    state = build_state(dev, kind=wl["kind"], out=wl["out"])
    n_slots = 4
    eng = EvalEngine(state, batch, (SRC_H, SRC_W), wl["out"], device=dev, n_slots=n_slots)
    # ---- device-resident inputs: n_slots distinct batches (4 x 207 MB > 126 MB L2).  Weak-scaling workloads: every
    #      rank has its own pixels and owns the logical index range [r*steps*batch, (r+1)*steps*batch).  Sharded
    #      workload: the pixels of logical image i are slot (i // batch) % n_slots, row i % batch of ONE ring that is
    #      the same on every rank, and the job's global batches are dealt to the ranks in contiguous ranges.
    ring = device_u8_batches(n_slots, batch, SRC_H, SRC_W, seed=100 + (0 if sharded else rank), device=dev)
    for s in range(n_slots):
        eng.u8[s].copy_(ring[s])
    del ring
    if sharded:
        n_images = int(args.images or wl["images"])
        b_lo, b_hi, n_batches = D.shard_batches(n_images, batch, rank, world)
        steps = b_hi - b_lo
        step_batches = [b_lo] * warmup + list(range(b_lo, b_hi))          # warm-up repeats the first batch (un-counted)
    else:
        n_images = None
        steps = args.steps
        base = rank * steps
        step_batches = [base] * warmup + [base + s for s in range(steps)]
    total_steps = warmup + steps

    def batch_meta(b):
        """labels / group ids of global batch b; a ragged last batch is padded with group id 255 = counted nowhere"""
        idx = np.arange(batch, dtype=np.int64) + b * batch
        label, ftype, sex, control = counter_metadata(idx, seed=7)
        grp = np.stack([ftype, sex, control])
        if n_images is not None and (b + 1) * batch > n_images:
            grp[:, max(0, n_images - b * batch):] = 255
        return label, grp

    meta = [batch_meta(b) for b in step_batches]
    lab_all = torch.from_numpy(np.stack([m[0] for m in meta])).to(dev)
    grp_all = torch.from_numpy(np.stack([m[1] for m in meta])).to(dev)

    def resident_step(s):
        slot = step_batches[s] % n_slots
        with torch.cuda.stream(eng.stream):
            eng.label[slot].copy_(lab_all[s], non_blocking=True)
            eng.groups[slot].copy_(grp_all[s], non_blocking=True)
        eng.step_resident(slot)

    def barrier():
        # Local GPU work first, THEN the collective: an NCCL kernel that spin-waits for a slower peer must not share
        # the GPU with seconds of queued evaluation kernels (observed at N >= 2 with a 2 s region: a conv1 CTA made no
        # progress while the barrier kernel of an early rank was resident, until its mbarrier watchdog fired).
        if os.environ.get("SIA_BENCH_BARRIER_NOSYNC") != "1":      # (test switch: reproduce the co-residency case)
            torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    # ------------------------------------ value: inputs resident in HBM ------------------------------
    for s in range(warmup):
        resident_step(s)
    with torch.cuda.stream(eng.stream):
        # warm-up of the collective too: NCCL loads its int64-sum kernel and sets the stream up lazily on the first
        # call (about a millisecond -- 10 % of a 20-step timed region)
        D.allreduce_counts(torch.zeros_like(eng.counts))
    eng.synchronize()
    eng.reset_counts()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # Only rank 0 samples (it prints the line), and only after the region's launches have been issued (ClockSampler).
    # The sampler is set up BEFORE the barrier: NVML initialisation takes milliseconds, and a rank that enters the
    # region late makes every other rank wait for it in the closing all-reduce (measured: +45 % at 2 GPUs).
    clocks = ClockSampler(local, enabled=(rank == 0))
    barrier()
    with clocks:
        e0.record(eng.stream)
        for s in range(warmup, total_steps):
            resident_step(s)
        eng.allreduce_counts()                     # the job's single collective (576 bytes), on the engine stream
        e1.record(eng.stream)
        clocks.sample_until(e1)                    # every launch of the region is issued; the GPU is still running it
        barrier()
    dt = D.max_over_ranks(e0.elapsed_time(e1) * 1e-3, device=dev)
    images_done = n_images if sharded else world * steps * batch
    value = images_done / dt
    counts_resident = eng.read_counts()

    # ------------------------------------ configs[2] check: 1-GPU recount, outside every timed region (and before the e2e
    #      phase, which refills the input slots from two host buffers) ---
    recount = None
    if sharded and rank == 0:
        eng.reset_counts()
        all_meta_t0 = time.perf_counter()
        for b in range(n_batches):
            label, grp = batch_meta(b)
            slot = b % n_slots
            with torch.cuda.stream(eng.stream):
                eng.label[slot].copy_(torch.from_numpy(label).to(dev, non_blocking=False), non_blocking=True)
                eng.groups[slot].copy_(torch.from_numpy(grp).to(dev, non_blocking=False), non_blocking=True)
            eng.step_resident(slot)
        single = eng.read_counts()
        recount = {"images": int(single[0].sum()), "seconds_incl_host_metadata": time.perf_counter() - all_meta_t0,
                   "bit_identical_to_allreduced": bool(torch.equal(single, counts_resident))}
        assert recount["bit_identical_to_allreduced"], "all-reduced counts differ from the 1-GPU recount"
        assert recount["images"] == n_images

    # ------------------------------------ sustained: >= 2 s of back-to-back steps, own clock record --------------
    sustained = None
    if not sharded and not args.no_sustained:
        n_sus = max(steps, int(args.sustained_seconds / max(dt / steps, 1e-6)))
        sus_clocks = ClockSampler(local, enabled=(rank == 0))
        eng.reset_counts()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        with sus_clocks:
            s0.record(eng.stream)
            for k in range(n_sus):
                eng.step_resident(k % n_slots)
            s1.record(eng.stream)
            sus_clocks.sample_until(s1)
            barrier()
        dts = D.max_over_ranks(s0.elapsed_time(s1) * 1e-3, device=dev)
        sustained = {"value": world * n_sus * batch / dts, "unit": UNIT, "steps": n_sus, "seconds": dts,
                     "ms_per_step": 1e3 * dts / n_sus, "clocks": sus_clocks.summary()}
        eng.reset_counts()

    # ------------------------------------ e2e: pinned host buffers, public API ------------------------
    host_u8 = [torch.empty((batch, SRC_H, SRC_W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
    for i, h in enumerate(host_u8):
        h.copy_(eng.u8[i].cpu())
    e2e_steps = min(steps, args.e2e_steps) if sharded else steps
    host_lab = [torch.from_numpy(m[0]).pin_memory() for m in meta[:warmup + e2e_steps]]
    host_grp = [torch.from_numpy(m[1]).pin_memory() for m in meta[:warmup + e2e_steps]]
    host_counts = torch.zeros_like(eng.counts, device="cpu").pin_memory()
    h2d = batch * SRC_H * SRC_W * 3 + batch + N_ATTR * batch
    d2h = host_counts.numel() * 8
    # what the e2e figure is held against: a plain pinned copy of one batch, (a) this GPU alone -- the ranks take turns --
    # and (b) every rank at once (GPUs of one box share host uplinks, so (b) per GPU can be far below (a))
    h2d_alone = 0.0
    for r in range(world):
        if r == rank:
            h2d_alone = pinned_h2d_peak(dev, batch * SRC_H * SRC_W * 3)
        barrier()
    h2d_together = pinned_h2d_peak(dev, batch * SRC_H * SRC_W * 3) if world > 1 else h2d_alone
    h2d_alone_min = -D.max_over_ranks(-h2d_alone, device=dev)
    h2d_together_min = -D.max_over_ranks(-h2d_together, device=dev)
    for s in range(min(warmup, 3)):
        eng.step(host_u8[s % 2], host_lab[s], host_grp[s], slot=s % n_slots)
    eng.synchronize()
    eng.reset_counts()
    barrier()
    t0 = time.perf_counter()
    for s in range(warmup, warmup + e2e_steps):
        eng.step(host_u8[s % 2], host_lab[s], host_grp[s], slot=s % n_slots)
        with torch.cuda.stream(eng.stream):
            host_counts.copy_(eng.counts, non_blocking=True)       # the step's metric, D2H
    eng.allreduce_counts()
    barrier()
    dt_e2e = D.max_over_ranks(time.perf_counter() - t0, device=dev)
    e2e_images = world * e2e_steps * batch
    e2e_value = e2e_images / dt_e2e
    counts_e2e = eng.read_counts()
    e2e_gbs_per_gpu = h2d * e2e_steps / dt_e2e / 1e9

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # ------------------------------------ rank 0: roofline, CPU baseline, report ----------------------
    clock_summary = clocks.summary()
    # the per-stage times are taken in a short burst of their own (12 iterations): the GEMM figure they are held against
    # follows from the clocks of THAT burst, not from the timed region's (a 1 M-image region runs power-capped, the
    # stage burst right after it may not)
    stage_clocks = ClockSampler(local, enabled=True)
    with stage_clocks:
        stage_done = torch.cuda.Event()
        stages = stage_breakdown(eng, batch, peaks, None, done_event=stage_done, sampler=stage_clocks)
    stage_clock_summary = stage_clocks.summary()
    tensor_tflops, tensor_note = tensor_peak(peaks, stage_clock_summary)
    for v in stages.values():
        if v.get("bound") == "tensor":
            v["peak"], v["frac"] = tensor_tflops, v["achieved"] / tensor_tflops
            if v["frac"] > 1.0:
                # the measured denominator is a cuBLAS bf16 GEMM on this pool's parts, not a hardware ceiling: a kernel
                # that beats it is held against the nominal dense peak (2.25 PFLOP/s at the maximum SM clock) instead
                mx = stage_clock_summary.get("sm_max_mhz") or 1965.0
                sm = stage_clock_summary.get("sm_mhz") or mx
                v["peak_measured"], v["frac_of_measured"] = v["peak"], v["frac"]
                v["peak"] = NOMINAL_BF16_TFLOPS * sm / mx
                v["frac"] = v["achieved"] / v["peak"]
                v["peak_note"] = ("achieved exceeds the cuBLAS-measured bf16 GEMM figure; held against the nominal dense "
                                  "bf16 peak scaled to the observed SM clock")
    dominant = max((k for k in stages if "bound" in stages[k]), key=lambda k: stages[k]["ms"])
    d = stages[dominant]
    step_ms = 1e3 * dt / steps
    roofline = {"kernel": dominant, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"],
                "unit": d["unit"], "frac": d["frac"], "traffic": ncu_traffic(dominant, batch),
                "peak_source": peaks["source"] + (": " + tensor_note if d["bound"] == "tensor" else ": HBM copy GB/s"),
                "kernel_share_of_step": d["ms"] / step_ms,
                "whole_step_tflops": sum(v.get("flops_per_image", 0) for v in stages.values()) * batch / (dt / steps) / 1e12}
    cpu = cpu_baseline() if (world == 1 and not args.no_cpu_baseline and args.workload == "list224") else None
    summary = tt.results_from_type_counts(counts_resident, out=lambda *a: None)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(wl, batch, world, n_slots),
        "run": {"cuda_graph": True, "host_numa_node_rank0": numa["numa_node"], "input_slots": n_slots},
        "clocks": clock_summary, "stage_clocks": stage_clock_summary,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * dt_e2e / e2e_steps, "steps": e2e_steps,
                "roofline": {"bound": "pcie_h2d", "achieved": e2e_gbs_per_gpu, "unit": "GB/s per GPU",
                             "peak": h2d_together_min, "frac": e2e_gbs_per_gpu / h2d_together_min,
                             "peak_alone": h2d_alone_min,
                             "note": "peak = measured pinned-host -> device copy of one batch with all ranks copying at "
                                     "once (min over ranks); peak_alone = the same copy with one rank at a time"}},
        "sustained": sustained,
        "gpu_launches": steps * eng.launches_per_batch,
        "roofline": roofline, "stages": stages, "cpu_baseline": cpu,
        "check": {"images_counted": int(counts_resident[0].sum()), "expected": images_done,
                  "e2e_images_counted": int(counts_e2e[0].sum()), "accuracy": summary["accuracy"],
                  "tone_di": summary["tone_di_results"]["di"], "recount_1gpu": recount},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU per step (default: the workload's)")
    ap.add_argument("--workload", default="list224", choices=sorted(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="images per step of the reference arm (default: the workload's 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained sub-record")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--images", type=int, default=0, help="logical images of the sharded workload (default 1M)")
    ap.add_argument("--e2e-steps", type=int, default=100, help="e2e steps of the sharded workload")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return
    try:
        run_ours(args)
    except Exception:
        try:        # a kernel watchdog (mbarrier wait timed out -> trap) leaves its site in pinned host memory
            from skin_image_analysis_b200 import _lib
            wd = _lib.load().sia_watchdog_status(0)
            print(f"[bench] sia watchdog word 0x{wd:08x} (site {(wd >> 16) & 0x7fff}, block {wd & 0xffff})",
                  file=sys.stderr, flush=True)
        except Exception:
            pass
        raise


if __name__ == "__main__":
    main()
