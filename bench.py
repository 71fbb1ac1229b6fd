#!/usr/bin/env python
"""Benchmark of the MPNN forward hot path (BASELINE.json metric: ion-pair graphs/s, viscosity model).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One "step" = one forward of the whole viscosity MPNN over one packed batch of synthetic ion pairs.
Workload at N GPUs (weak scaling): BASELINE.json configs[2] -- the viscosity inference sweep -- with
``--pairs-per-gpu`` pairs resident on every GPU (default 2,097,152 => 16.8 M pairs at 8 GPUs, the config's 16 M).
Pairs are independent, so ranks shard them with no data-path collective.

value   = pairs processed by all ranks / max-over-ranks device time, inputs already resident in HBM.
e2e     = same metric through MPNNModel.predict-style calls from PINNED HOST buffers: per step the packed
          batch is copied host->device, the kernels run, and the predictions are copied device->host.
roofline= the dominant kernel (largest share of the step), algorithmic bytes per launch / its CUDA-event time,
          against MEASURED_PEAKS.json.
cpu_baseline = oracle/ref_model.py (torch fp32 port of the reference's TF graph; TensorFlow is not installable
          here) on the host cores, on a bounded sample of the same workload, rank 0 only.
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

METRIC = "ion_pair_graphs_per_s"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="visc_sweep", choices=["visc_sweep", "mp64k", "visc_train", "wide"],
                    help="visc_sweep = BASELINE configs[2] (headline); mp64k = configs[1]; visc_train = configs[3] "
                         "(forward + backward + all-reduce + Adam, 65,536 pairs per GPU)")
    ap.add_argument("--pairs-per-gpu", type=int, default=None)
    ap.add_argument("--precision", default="fp16", choices=["fp32", "bf16", "bf16_precise", "fp16", "fp16_precise"],
                    help="fp16 / bf16 = tcgen05 tensor-core path (16-bit operands, fp32 accumulate; the 2e-2 path, default); "
                         "fp32 = SIMT 1e-5 parity path")
    ap.add_argument("--staged", action="store_true", help="per-layer kernels instead of the fused whole-tower kernel")
    ap.add_argument("--tc-flags", type=int, default=0, help="extra IMP_TC_* tuning flags (e.g. 4 = MP8)")
    ap.add_argument("--skewed", action="store_true", help="Zipf(1.2) bond types instead of uniform")
    ap.add_argument("--cpu-sample-pairs", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sync-every-step", action="store_true", help="visc_train: read the loss back after every step")
    ap.add_argument("--e2e-chunks", type=int, default=4, help="host-resident chunks the e2e leg streams per step")
    ap.add_argument("--no-extras", action="store_true",
                    help="default run only: skip the sub-records of the other BASELINE configs (extra.mp64k / train / wide / fp32)")
    ap.add_argument("--min-timed-s", type=float, default=0.0,
                    help="raise --steps so that the timed region lasts at least this long (sub-records use 1.2 s)")
    return ap.parse_args()


def workload_config(args):
    if args.workload == "visc_sweep":
        P = args.pairs_per_gpu or 2_097_152
        kind = "viscosity"
        name = ("BASELINE configs[2]: viscosity MPNN inference sweep (atom_dim 32, bond_dim 8, 4 steps, vocab 123/71, "
                "10-40 atoms/ion), weak scaling")
    elif args.workload == "wide":
        P = args.pairs_per_gpu or 16_384
        kind = "viscosity"
        name = ("BASELINE configs[4]: wide/deep viscosity MPNN forward (atom_dim 256, bond_dim 8, 6 steps, 40-120 atoms/ion); "
                "fp16 = tcgen05 GEMM kernels (csrc/wide_tc.cu), fp32 = general-shape SIMT kernels")
    elif args.workload == "visc_train":
        P = args.pairs_per_gpu or 65_536
        kind = "viscosity"
        name = ("BASELINE configs[3]: viscosity MPNN training step (forward + backward, NCCL gradient all-reduce, "
                "per-variable clipnorm + Adam), 65,536 pairs per GPU")
    else:
        P = args.pairs_per_gpu or 65_536
        kind = "melting_point"
        name = "BASELINE configs[1]: melting-point MPNN forward, 64k pairs per GPU (atom_dim 32, bond_dim 1024, 4 steps)"
    return P, kind, name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic(kernel, pairs):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of the SAME launch size
    (profiles/r01_fused_traffic.json); None when there is no capture for this size."""
    for name in ("r02_fused_traffic.json", "r01_fused_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            d = json.load(open(p))
            if d.get("kernel") == kernel and d.get("pairs_per_launch") == pairs:
                return d["dram_bytes_read"] + d["dram_bytes_write"], d["source"]
    return None, None


def simt_pipe_ceiling(batch, S, ms, sm_mhz, n_sm=148):
    """The fused forward's second ceiling: its SIMT side.  Pipe-time model of one launch with the pipe rates
    measured on B200 by tools/issue_microbench.cu (warp-instructions per clock per SM sub-partition): HFMA2 0.5, fp32
    FFMA / FADD / FMUL 1.0, MUFU.TANH 0.125.
      FMA pipe:  Z build 128 HFMA2 per unique entry and step (2 cycles each, at 100 % lane use: / 32 lanes)
                 + 352 fp32 instructions per atom and step (gates 96, r*h 32, blend 64, statistics 64, LayerNorm 96)
      XU pipe:   97 MUFU per atom and step (z, r, candidate tanh; rsqrt), 8 cycles each
    Ceiling = the busier pipe 100 % busy on every sub-partition of every SM."""
    N, Eu = batch.n_atoms, batch.n_unique
    fma = S * (2.0 * 128 * Eu / 32 + 352.0 * N / 32)
    xu = S * 8.0 * 97 * N / 32
    smsp_cycles_per_s = n_sm * 4 * (sm_mhz or 1965.0) * 1e6
    t_ceiling = max(fma, xu) / smsp_cycles_per_s
    return {"model": "FMA pipe: S (8 Eu + 11 N) cycles, XU pipe: S 24.25 N cycles per launch and sub-partition-sum; rates "
                     "measured by tools/issue_microbench.cu (HFMA2 0.5, fp32 1.0, MUFU 0.125 warp-instr/clk/SMSP)",
            "fma_pipe_cycles": fma, "xu_pipe_cycles": xu, "pairs_per_s_at_ceiling": batch.n_pairs / t_ceiling,
            "frac": t_ceiling / (ms * 1e-3), "fma_pipe_busy": fma / smsp_cycles_per_s / (ms * 1e-3),
            "xu_pipe_busy": xu / smsp_cycles_per_s / (ms * 1e-3)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops_sustained", d.get("bf16_tflops")), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ reference arm / cpu baseline
def cpu_reference_run(kind, n_pairs, steps, warmup, skewed, wide=False):
    """Times oracle/ref_model.py (fp32, all host threads) the way the reference predicts: padded inputs,
    ``model.predict(x)`` with Keras' default batch size 32 (train_viscosity.py:366)."""
    import torch

    from ionic_mpnn_b200 import synth
    from oracle import ref_inputs, ref_model

    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    if wide:  # configs[4]: ~240x the arithmetic per pair, and a (B, E, d, d) temporary of 68 MB per pair
        n_pairs = min(n_pairs, 64)
    recs = synth.make_records(n_pairs, seed=1002, skewed=skewed, label="log_eta" if kind == "viscosity" else "mp",
                              **({"n_min": 40, "n_max": 120} if wide else {}))
    spec = ref_model.make_spec(kind, atom_dim=256, num_steps=6) if wide else ref_model.make_spec(kind)
    params = ref_model.init_params(spec, seed=1)
    x = ref_inputs.build_inputs(recs, with_temperature=kind == "viscosity")
    best = None
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        ref_model.predict(spec, params, x, dtype=torch.float32, batch_size=32)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    mean = sum(times) / len(times)
    # the same port with a large batch ("also batch_size=1024 for a fairer number"), one run after the warm-up
    big = None
    if not wide:
        t0 = time.perf_counter()
        ref_model.predict(spec, params, x, dtype=torch.float32, batch_size=1024)
        big = n_pairs / (time.perf_counter() - t0)
    return {"value": n_pairs / mean, "value_batch1024": big, "unit": UNIT, "cores": cores, "kind": "port", "n_pairs": n_pairs,
            "sample": f"{n_pairs} synthetic pairs (seed 1002), padded as the reference pads, predict(batch_size=32), "
                      f"torch fp32 CPU port of models/layers.py, mean of {len(times)} runs after {warmup} warm-up",
            "ms_per_step": mean * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    P, kind, name = workload_config(args)
    warm = max(1, min(args.warmup, 2))
    steps = max(1, min(args.steps, 5))
    r = cpu_reference_run(kind, args.cpu_sample_pairs, steps, warm, args.skewed, wide=args.workload == "wide")
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "sample_pairs_per_step": r["n_pairs"]},
            "cpu_baseline": {k: r[k] for k in ("value", "value_batch1024", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def props_sm_count():
    import torch

    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count


def stage_flops(batch, d, S):
    """Algorithmic FLOP per launch: messages 2*E*d^2 with E = live entries counted with multiplicity,
    gated update 12*N*d^2, per step."""
    N, E = batch.n_atoms, batch.n_edges
    return {"mpnn_forward_fused": S * (2 * E + 12 * N) * d * d, "mpnn_forward_fused_planned": S * (2 * E + 12 * N) * d * d,
            "gated_update_tc": 12 * N * d * d,
            "gated_update_wide": 12 * N * d * d, "message_agg": 2 * batch.n_unique * d * d,
            # wide tensor path: algorithmic message work is 2*E*d^2 (the kernel executes 16*N*d^2 as Z.Wc, K = 8d)
            "wide_message": 2 * E * d * d, "wide_gated_update": 12 * N * d * d,
            "edge_messages_tc": 2 * batch.n_unique * d * d, "reduce_gated_update_tc": 12 * N * d * d,
            "edge_messages_tc16": 2 * batch.n_unique * d * d, "reduce_gated_update_tc16": 12 * N * d * d,
            "edge_messages_tc16_planned": 2 * batch.n_unique * d * d,
            "gated_update": 12 * N * d * d, "gated_update_tc32": 12 * N * d * d,
            "edge_messages_grouped": 2 * batch.n_unique * d * d, "edge_messages_grouped_tc32_planned": 2 * batch.n_unique * d * d}


def stage_bytes(batch, d, S, s=4):
    """Algorithmic bytes per launch (each operand once, weights / tables amortised to 0, int32 indices)."""
    N, Eu, P = batch.n_atoms, batch.n_unique, batch.n_pairs
    return {
        # fused forward: only the index stream is read (atom_id, row_ptr, mol_ptr, (src, bond|mult) per entry) and
        # the molecule sums are written (SURVEY 8d "stretch": bytes -> indices only)
        "mpnn_forward_fused": 8 * N + 8 * Eu + 8 * P + 2 * P * d * 4,
        # planned forward: one 2 KiB record per 128-row tile (~N / 125 tiles) in, the molecule sums out; the plan kernel reads
        # the index stream once and writes those records
        "mpnn_forward_fused_planned": (N // 125 + 1) * 2048 + 2 * P * d * 4,
        "fused_plan": 8 * N + 8 * Eu + 8 * P + (N // 125 + 1) * 2048,
        "readout_visc": 2 * P * d * 4 + 8 * P,
        "readout_mp": 2 * P * d * 4 + 4 * P,
        "global_sum_pool": N * d * 4 + 4 * N + 2 * P * d * 4,
        "embed_atoms": 4 * N + N * d * s,
        "message_agg": 2 * N * d * s + 8 * Eu + 4 * N,   # h in, agg out, (src, bond|mult) per unique entry, row_ptr
        # grouped tcgen05 message GEMM (csrc/msg_tc.cu): one gathered source row in, one message row out per unique
        # entry, bucket_perm + src + bond|mult; then the CSR segment sum
        "edge_messages_tc": Eu * (2 * d * 4 + 12),
        "segment_sum": Eu * d * 4 + N * d * 4 + 4 * N,
        "reduce_gated_update_tc": Eu * d * 4 + 4 * N + 2 * N * d * 4,   # message rows + row_ptr + h in / h out (agg never written)
        # 16-bit I/O forms: gathered h16 row + 16-bit message row + indices; 16-bit messages + row_ptr + h in / h out + h16 out
        "edge_messages_tc16": Eu * (2 * d * 2 + 12),
        "edge_messages_tc16_planned": Eu * (2 * d * 2 + 12),
        "edge_messages_tc16_plan": Eu * (4 + 8 + 8),   # bucket_perm in; src, bond|mult gathered; bucket-ordered copies out
        "reduce_gated_update_tc16": Eu * d * 2 + 4 * N + 2 * N * d * 4 + N * d * 2,
        "embed_atoms16": 4 * N + N * d * 6,
        "gated_update": 3 * N * d * s,                    # h, agg in; h out
        "gated_update_tc32": 3 * N * d * s,
        "edge_messages_grouped": Eu * (2 * d * 4 + 12), "edge_messages_grouped_tc32_planned": Eu * (2 * d * 4 + 12),
        "gated_update_tc": 3 * N * d * s,
        "gated_update_wide": 3 * N * d * s,
        # wide tensor path (16-bit operand copies next to the fp32 state, csrc/wide_tc.cu)
        "wide_embed": 4 * N + N * d * 6,
        "wide_message": N * d * 2 + N * d * 2 + 8 * Eu + 4 * N,          # h16 in, agg16 out, entries, row_ptr
        "wide_gated_update": N * d * (2 + 2 + 4) + N * d * (4 + 2),      # h16, agg16, h32 in; h32, h16 out
        "wide_pool": N * d * 4 + 4 * N + 8 * P + 2 * P * d * 4,
        "pool_head": N * d * s + 4 * N + 8 * P + 8 * P,   # h, atom_id, mol_ptr (2 towers), T + out
    }


def run_train(args):
    """configs[3]: one step = loss_and_grads + one flat all-reduce + clip/Adam on a resident batch (weak scaling)."""
    import torch
    import torch.distributed as dist

    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    P, kind, name = workload_config(args)
    spec = make_spec(kind)
    model = MPNNModel(spec, device=f"cuda:{local}", seed=0, precision="fp32")
    batch, _, _ = graph.synth_batch(P, seed=2003 + rank, skewed=args.skewed)
    batch.target = __import__("numpy").random.default_rng(7 + rank).normal(2.0, 1.0, size=P).astype("float32")
    batch.to(f"cuda:{local}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(args.warmup):
        losses.append(model.train_step(batch))
    barrier()
    args.steps = steps_for_min_time(args, lambda: model.train_step(batch), world, local)
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
        time.sleep(0.25)
    barrier()  # the step contains a collective: every rank must enter the timed region together
    t0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        losses.append(model.train_step(batch))
        if args.sync_every_step:
            losses[-1].item()  # what a training loop that logs its loss does
    ev1.record()
    t_enq = time.time() - t0
    torch.cuda.synchronize()
    t1 = time.time()
    barrier()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    # per-kernel share: one more instrumented step.  EVERY rank runs it (the step contains the all-reduce);
    # rank 0 reports its own events.
    per_kernel = {}
    from ionic_mpnn_b200 import _lib
    import ionic_mpnn_b200.train as tt
    real_call = _lib.call
    events = []

    def timed_call(n, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_call(n, *a)
        e1.record()
        events.append((n, e0, e1))

    real_ar = dist.all_reduce

    def timed_all_reduce(t, *a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_ar(t, *a, **k)
        e1.record()
        events.append(("nccl_all_reduce", e0, e1))

    tt._lib.call = timed_call
    dist.all_reduce = timed_all_reduce
    try:
        model.train_step(batch)
    finally:
        tt._lib.call = real_call
        dist.all_reduce = real_ar
    torch.cuda.synchronize()
    for n, e0, e1 in events:
        per_kernel[n.replace("imp_", "")] = per_kernel.get(n.replace("imp_", ""), 0.0) + e0.elapsed_time(e1)
    barrier()
    if rank == 0:
        lv = [float(l.item()) for l in losses]
        tot = sum(per_kernel.values()) or 1.0
        # ---- roofline of the dominant kernel.  Algorithmic work per launch: every training kernel at
        # atom_dim 32 is to the left of the ridge (<= 58 FLOP per byte against 211), so the bound is HBM.
        N, Eu, d, S = batch.n_atoms, batch.n_unique, 32, spec["num_steps"]
        tb = {"gated_update_bwd": 5 * N * d * 4, "gated_update": 3 * N * d * 4, "edge_messages_grouped": Eu * (2 * d * 4 + 12), "edge_messages_grouped_tc32": Eu * (2 * d * 4 + 12), "edge_messages_grouped_tc32_planned": Eu * (2 * d * 4 + 12),
              # stored-gate forms: the forward also writes z, r, tanh(.); the backward reads h, agg, z, r, tanh(.), g_out and
              # writes dh, dagg
              "gated_update_train": 6 * N * d * 4, "gated_update_tc32": 6 * N * d * 4, "gated_update_bwd_stored": 8 * N * d * 4, "gated_update_bwd_tc": 8 * N * d * 4,
              "segment_sum": Eu * d * 4 + N * d * 4 + 4 * N, "segment_sum_add": Eu * d * 4 + 2 * N * d * 4 + 4 * N,
              "bond_transform_bwd": Eu * (2 * d * 4 + 12), "embed_atoms": 4 * N + N * d * 4, "embed_bwd": 4 * N + N * d * 4,
              "bond_occurrence_norm2": S * Eu * 2 * d * 4 + 12 * Eu, "sumsq": N * d * 4,
              "global_sum_pool": N * d * 4 + 4 * N, "pool_bwd": N * d * 4 + 4 * N}
        tf = {"gated_update_bwd": 36 * N * d * d, "gated_update": 12 * N * d * d, "edge_messages_grouped": 2 * Eu * d * d, "edge_messages_grouped_tc32": 2 * Eu * d * d, "edge_messages_grouped_tc32_planned": 2 * Eu * d * d,
              # without recomputation: 3 input-gradient + 3 weight-gradient contractions of 2 * 64 * 32 FLOP per atom (the
              # tcgen05 kernel executes each as three tf32 MMAs)
              "gated_update_train": 12 * N * d * d, "gated_update_tc32": 12 * N * d * d, "gated_update_bwd_stored": 24 * N * d * d, "gated_update_bwd_tc": 24 * N * d * d,
              "bond_transform_bwd": 2 * Eu * d * d, "bond_occurrence_norm2": S * 2 * Eu * d * d * 8}
        counts = {}
        for n, _, _ in events:
            counts[n.replace("imp_", "")] = counts.get(n.replace("imp_", ""), 0) + 1
        hbm_peak, tf_peak, peak_src = measured_peaks()
        pk = {k: {"avg_ms": v / counts[k], "launches_per_step": counts[k], "share": v / tot,
                  "GBps": tb[k] / (v / counts[k] * 1e-3) / 1e9 if k in tb else None,
                  "TFLOPs": tf[k] / (v / counts[k] * 1e-3) / 1e12 if k in tf else None} for k, v in per_kernel.items()}
        dom = max((k for k in per_kernel if k in tb), key=lambda k: per_kernel[k])
        ach = pk[dom]["GBps"]
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": tb[dom],
                    "algorithmic_flop_per_launch": tf.get(dom), "avg_launch_ms": pk[dom]["avg_ms"],
                    "share_of_step": pk[dom]["share"], "per_kernel": pk}
        line = {"metric": "train_ion_pair_graphs_per_s", "value": P * world * args.steps / (ms_total * 1e-3), "unit": UNIT,
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": name, "pairs_per_gpu": P, "global_batch": P * world, "atoms_per_gpu": batch.n_atoms,
                           "params": model.count_params(), "collective": "one NCCL all-reduce of the flat gradient bucket "
                           f"({model._train['grad'].numel()} fp32) per step" if world > 1 else "none (1 GPU)",
                           "l2_policy": "inputs larger than L2 (%.1f GB of saved activations per GPU)" % (
                               9 * batch.n_atoms * 32 * 4 / 1e9)},
                "gpu_launches": sum(1 for e in events if e[0] != "nccl_all_reduce") * args.steps, "clocks": clocks,
                "roofline": roofline,
                "all_reduce": {"per_step": counts.get("nccl_all_reduce", 0), "ms_per_step": round(per_kernel.get("nccl_all_reduce", 0.0), 4),
                               "floats": int(model._train["grad"].numel()),
                               "carries": "gradients + [sse, pair count, 2 occurrence norms]"},
                "embedding_clip": "per-occurrence norm (Keras IndexedSlices semantics)",
                "host_enqueue_ms_per_step": round(t_enq * 1e3 / args.steps, 2), "loss_first_last": [lv[0], lv[-1]], "loss_decreased": lv[-1] < lv[0],
                "kernel_ms": {k: round(v, 3) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
                "kernel_share": {k: round(v / tot, 3) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])}}
        return line
    return None


def steps_for_min_time(args, one_step, world, local):
    """--min-timed-s: number of timed steps such that the region lasts at least that long (the clock sampler needs ~1 s);
    the same on every rank (max over ranks of one probe step)."""
    import torch
    import torch.distributed as dist

    if not args.min_timed_s:
        return args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    one_step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return max(args.steps, int(args.min_timed_s * 1e3 / max(float(t.item()), 1e-3)) + 1)


def run_b200(args):
    import torch
    import torch.distributed as dist

    from ionic_mpnn_b200 import graph
    from ionic_mpnn_b200.model import MPNNModel, make_spec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    P, kind, name = workload_config(args)
    wide = args.workload == "wide"
    spec = make_spec(kind, atom_dim=256, num_steps=6) if wide else make_spec(kind)
    if wide and args.precision not in ("fp16", "fp16_precise"):
        args.precision = "fp32"
    model = MPNNModel(spec, device=f"cuda:{local}", seed=0, precision=args.precision,
                      fused=False if args.staged else "auto")
    model.extra_tc_flags = args.tc_flags
    model.fp32_tensor = bool(getattr(args, "fp32_tensor", False))
    d, S = spec["atom_dim"], spec["num_steps"]

    t_pack0 = time.perf_counter()
    nmin, nmax = (40, 120) if wide else (10, 40)
    batch, _, _ = graph.synth_batch(P, seed=1003 + rank, n_min=nmin, n_max=nmax, skewed=args.skewed,
                                    with_temperature=(kind == "viscosity"))
    t_pack = time.perf_counter() - t_pack0
    batch.to(f"cuda:{local}")
    model.refresh_tables()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        model.forward_packed(batch)
    barrier()
    args.steps = steps_for_min_time(args, lambda: model.forward_packed(batch), world, local)
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
        time.sleep(0.25)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        out = model.forward_packed(batch)
    ev1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = P * world * args.steps / (ms_total * 1e-3)

    # ---- per-kernel timing (rank 0): CUDA events around every launch of one more pass -----------
    per_kernel = {}
    if rank == 0:
        import ctypes as C

        from ionic_mpnn_b200 import _lib
        real_call = _lib.call
        events = []

        def timed_call(name, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            real_call(name, *a)
            e1.record()
            events.append((name, e0, e1))

        import ionic_mpnn_b200.model as mm
        mm._lib.call = timed_call
        model.wide_per_stage_calls = True  # wide path: one ABI call per kernel, so that each one is timed
        try:
            for _ in range(2):
                model.forward_packed(batch)
        finally:
            mm._lib.call = real_call
            model.wide_per_stage_calls = False
        torch.cuda.synchronize()
        for kname, e0, e1 in events:
            per_kernel.setdefault(kname.replace("imp_", ""), []).append(e0.elapsed_time(e1))
    barrier()

    # ---- end-to-end from pinned host buffers --------------------------------------------------
    # The public streaming call (MPNNModel.predict_stream): the workload as host-resident packed chunks; per step every
    # chunk is copied host->device from pinned memory (copy stream, double-buffered against the kernels), the kernels
    # run, and the predictions are copied device->host into a pinned result vector.
    e2e = None
    if not args.no_e2e:
        n_chunks = max(1, min(args.e2e_chunks, P // 1024)) if P >= 2048 else 1
        per = P // n_chunks
        del batch.dev
        batch.dev = None
        torch.cuda.empty_cache()
        chunks = []
        for c in range(n_chunks):
            pc = per if c + 1 < n_chunks else P - per * (n_chunks - 1)
            ch, _, _ = graph.synth_batch(pc, seed=5003 + 97 * rank + c, n_min=nmin, n_max=nmax, skewed=args.skewed,
                                         with_temperature=(kind == "viscosity"))
            chunks.append(ch.pin())
        out_host = torch.empty(P, dtype=torch.float32).pin_memory()
        h2d = 0
        for _ in range(max(1, args.warmup - 1)):
            _, h2d = model.predict_stream(chunks, out_host)
        barrier()
        ev0.record()
        for _ in range(args.steps):
            model.predict_stream(chunks, out_host)
        ev1.record()
        torch.cuda.synchronize()
        model.check_status()
        barrier()
        ms2 = torch.tensor([ev0.elapsed_time(ev1)], device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e = {"value": P * world * args.steps / (float(ms2.item()) * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": P * 4, "ms_per_step": float(ms2.item()) / args.steps, "chunks_per_step": n_chunks,
               "input_feed": ("compact (16-bit atom words, 16-bit entry words)" if model._stream_state["fields"][3] == "edge_h"
                              else "compact (16-bit atom words, 32-bit entry words)") if model._stream_state["fields"][1] == "mol_eptr"
               else "int32 CSR arrays",
               "finite": bool(torch.isfinite(out_host).all()),
               "note": "MPNNModel.predict_stream: pinned host packed chunks -> H2D on a copy stream (2 staging slots) "
                       "-> kernels -> D2H predictions, every step, per GPU"}

    # ---- packers: the same flat ion arrays through imp_pack_host and imp_pack_device
    pack = None
    if rank == 0 and not args.no_e2e and not wide:
        pp = min(P, 262_144)
        cat_i = graph.synth_flat(pp, 9001, nmin, nmax, skewed=args.skewed)
        an_i = graph.synth_flat(pp, 9002, nmin, nmax, skewed=args.skewed)
        t0 = time.perf_counter()
        hb = graph.pack_flat(cat_i, an_i, spec["bond_vocab_size"])
        t_host = time.perf_counter() - t0
        Tc = batch.temperature[:pp] if batch.temperature is not None else None
        graph.pack_flat_device(cat_i, an_i, spec["bond_vocab_size"], device=f"cuda:{local}", temperature=Tc)  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            db = graph.pack_flat_device(cat_i, an_i, spec["bond_vocab_size"], device=f"cuda:{local}", temperature=Tc)
        torch.cuda.synchronize()
        t_dev = (time.perf_counter() - t0) / 3
        up = graph.upload_ions(cat_i, an_i, f"cuda:{local}")
        torch.cuda.synchronize()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(3):
            graph.pack_flat_device(cat_i, an_i, spec["bond_vocab_size"], device=f"cuda:{local}", uploaded=up)
        pe1.record()
        torch.cuda.synchronize()
        t_kern = pe0.elapsed_time(pe1) / 3 * 1e-3
        t0 = time.perf_counter()
        for _ in range(3):
            db = graph.pack_flat_device(cat_i, an_i, spec["bond_vocab_size"], device=f"cuda:{local}", temperature=Tc)
            o = model.forward_packed(db.as_compact() if model.compact_supported() and model.use_fused(db) else db)
            o_host = o.cpu()
        t_all = (time.perf_counter() - t0) / 3
        pack = {"pairs": pp, "host_packer_pairs_per_s": pp / t_host, "host_threads": len(os.sched_getaffinity(0)),
                "device_packer_pairs_per_s": pp / t_dev, "device_packer_resident_input_pairs_per_s": pp / t_kern,
                "ions_to_predictions_pairs_per_s": pp / t_all, "identical_n_unique": hb.n_unique == db.n_unique,
                "note": "device figures are wall-clock and include the H2D copy of the pageable flat ion arrays and the "
                        "host read-back of the counts; ions_to_predictions adds the fused forward and the D2H of the result"}

    if rank == 0:
        hbm_peak, tf_peak, peak_src = measured_peaks()
        sb = stage_bytes(batch, d, S)
        sf = stage_flops(batch, d, S)
        mean_ms = {k: sum(v) / len(v) for k, v in per_kernel.items()}
        tot_ms = {k: sum(v) / 2 for k, v in per_kernel.items()}  # two instrumented passes
        dom = max(tot_ms, key=tot_ms.get) if tot_ms else None
        roofline = None
        if dom:
            pk = {k: {"avg_ms": mean_ms[k], "launches_per_step": len(per_kernel[k]) // 2,
                      "share": tot_ms[k] / sum(tot_ms.values()),
                      "GBps": (sb[k] / (mean_ms[k] * 1e-3) / 1e9) if k in sb else None,
                      "TFLOPs": (sf[k] / (mean_ms[k] * 1e-3) / 1e12) if k in sf else None} for k in mean_ms}
            if dom in ("wide_message", "wide_gated_update"):
                # d = 256: 238+ FLOP per activation byte => tensor-bound; achieved = algorithmic FLOP of the
                # kernel / its launch time.  The message kernel executes 16*N*d^2 (Z.Wc with K = 8d) for 2*E*d^2
                # algorithmic; "executed_TFLOPs" reports what the tensor pipe actually ran.
                ach = sf[dom] / (mean_ms[dom] * 1e-3) / 1e12
                executed = {"wide_message": 16 * batch.n_atoms * d * d, "wide_gated_update": 12 * batch.n_atoms * d * d}
                for kk in executed:
                    if kk in pk:
                        pk[kk]["executed_TFLOPs"] = executed[kk] / (mean_ms[kk] * 1e-3) / 1e12
                step_flop = S * (2 * batch.n_edges + 12 * batch.n_atoms) * d * d
                roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                            "frac": ach / tf_peak, "traffic": None, "peak_source": peak_src + ", sustained bf16",
                            "algorithmic_flop_per_launch": sf[dom], "algorithmic_bytes_per_launch": sb.get(dom),
                            "executed_TFLOPs": executed[dom] / (mean_ms[dom] * 1e-3) / 1e12,
                            "whole_forward_algorithmic_TFLOPs": step_flop / (ms_total / args.steps * 1e-3) / 1e12,
                            "whole_forward_frac": step_flop / (ms_total / args.steps * 1e-3) / 1e12 / tf_peak}
            elif dom in ("mpnn_forward_fused", "mpnn_forward_fused_planned"):
                # the fused kernel keeps every activation on chip: 2.7 kFLOP per byte of index stream, far right of
                # the ridge (211 FLOP/B) => the tensor roofline is the one that bounds it
                ach = sf[dom] / (mean_ms[dom] * 1e-3) / 1e12
                traffic, traffic_src = ncu_traffic(dom, P)
                roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                            "frac": ach / tf_peak, "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": peak_src + ", sustained bf16",
                            "algorithmic_flop_per_launch": sf[dom], "algorithmic_bytes_per_launch": sb[dom],
                            "hbm_GBps_on_algorithmic_bytes": sb[dom] / (mean_ms[dom] * 1e-3) / 1e9,
                            "hbm_frac": sb[dom] / (mean_ms[dom] * 1e-3) / 1e9 / hbm_peak,
                            "issue_ceiling": simt_pipe_ceiling(batch, S, mean_ms[dom], (clocks or {}).get("sm_mhz"))}
            else:
                ach = sb.get(dom, 0) / (mean_ms[dom] * 1e-3) / 1e9
                roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": sb.get(dom)}
                if args.precision == "fp32" and dom in sf and not model.fp32_tensor:
                    # the exact fp32 kernels are SIMT: at atom_dim 32 the GatedUpdate has 32 FLOP per byte, to the RIGHT of the
                    # ridge of the fp32 FMA pipe (148 SMs x 128 lanes x 2 FLOP x clock / HBM peak = 11 FLOP per byte), so
                    # the ceiling that binds it is that pipe, not HBM: reported beside the HBM figure
                    mhz = (clocks or {}).get("sm_mhz") or 1965.0
                    fpeak = props_sm_count() * 128 * 2 * mhz * 1e6 / 1e12
                    roofline["fp32_pipe"] = {"achieved_TFLOPs": sf[dom] / (mean_ms[dom] * 1e-3) / 1e12, "peak_TFLOPs": fpeak,
                                             "frac": sf[dom] / (mean_ms[dom] * 1e-3) / 1e12 / fpeak,
                                             "peak_source": "SM count x 128 fp32 lanes x 2 x the SM clock sampled during the run"}
            roofline.update({"avg_launch_ms": mean_ms[dom], "share_of_step": tot_ms[dom] / sum(tot_ms.values()),
                             "per_kernel": pk})
        cpu = None
        if not args.no_cpu_baseline:
            r = cpu_reference_run(kind, args.cpu_sample_pairs, 3, 1, args.skewed, wide=wide)
            cpu = {k: r[k] for k in ("value", "value_batch1024", "unit", "cores", "kind", "sample")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16": "bf16", "bf16_precise": "bf16", "fp16": "f16",
                          "fp16_precise": "f16"}[args.precision] + ("" if args.precision == "fp32" else
                                                                    " operands, f32 accumulate, f32 atom states"),
                "data": "synthetic",
                "config": {"workload": name, "pairs_per_gpu": P, "atoms_per_gpu": batch.n_atoms,
                           "edges_per_gpu": batch.n_edges, "unique_edges_per_gpu": batch.n_unique,
                           "bond_types": "zipf1.2" if args.skewed else "uniform", "precision": args.precision,
                           "path": "wide tcgen05 GEMM kernels" if model.wide_supported() else
                           "fused whole-tower kernel" if model.use_fused(batch) else
                           "staged per-layer kernels" + (" (3xTF32 tensor-core GatedUpdate and messages)" if model.fp32_tensor else ""),
                           "l2_policy": "inputs larger than L2 (%.1f GB of indices%s per GPU)" % (
                               batch.nbytes() / 1e9, "" if model.use_fused(batch) else
                               " + %.1f GB of activations" % (3 * batch.n_atoms * d * 4 / 1e9)),
                           "parallelism": f"pairs sharded over {world} GPU(s), no collective",
                           "host_synth_and_pack_s": round(t_pack, 2)},
                "edges_per_s": batch.n_edges * world * args.steps / (ms_total * 1e-3),
                "gpu_launches": model.launches_per_forward(batch) * args.steps,
                "clocks": clocks, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "pack": pack}
        return line
    return None


EXTRAS = (  # (key, workload, precision): the other BASELINE configs as sub-records of the default line
    ("mp64k", "mp64k", "fp16"),        # configs[1]: melting-point forward, 64k pairs, staged tensor kernels
    ("train", "visc_train", "fp32"),   # configs[3]: training step (the only place a collective is timed)
    ("wide", "wide", "fp16"),          # configs[4]: atom_dim 256, 6 steps
    ("fp32", "visc_sweep", "fp32"),    # the 1e-5 parity path on the headline workload (exact fp32 SIMT kernels)
)


def run_extras(args):
    """Every rank runs every sub-workload (the training step contains the all-reduce); rank 0 collects the lines."""
    import copy
    import gc

    import torch

    out = {}
    for key, workload, precision in EXTRAS:
        a = copy.copy(args)
        a.workload, a.precision, a.pairs_per_gpu = workload, precision, None
        a.no_e2e = a.no_cpu_baseline = True
        a.staged, a.tc_flags, a.min_timed_s, a.steps, a.warmup = False, 0, 1.2, 3, 3
        if key in ("fp32", "fp32_tensor"):
            a.pairs_per_gpu = 524_288  # 8 M pairs/s: the full 2 M-pair batch would take 0.26 s per step
        a.fp32_tensor = key == "fp32_tensor"
        try:
            line = run_train(a) if workload == "visc_train" else run_b200(a)
        except Exception as e:  # a sub-record must never take the headline down with it
            line = {"error": f"{type(e).__name__}: {e}"} if int(os.environ.get("RANK", "0")) == 0 else None
        if line is not None:
            out[key] = line
        gc.collect()
        torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        line = run_train(args) if args.workload == "visc_train" else run_b200(args)
        headline = args.workload == "visc_sweep" and args.precision == "fp16" and not args.staged and not args.tc_flags
        if headline and not args.no_extras and args.pairs_per_gpu is None:
            extra = run_extras(args)
            if line is not None:
                line["extra"] = extra
        if line is not None:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
