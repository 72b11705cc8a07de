#!/usr/bin/env python
"""bench.py - user-item pairs/sec through TwoTowerBaseRetrieval.train_forward (+ backward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--d 128] [--batch 8192]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): TwoTowerBaseRetrieval, d = DU = DI = 128, F = IU = II = 128,
per-GPU batch 8192, hash tables 100 000 rows, T = 1, user_value_weights = [1.0]; synthetic seeded inputs,
reference default initialisation under torch.manual_seed(0).  One step = train_forward + loss.backward()
over one batch (8192 pairs per GPU).  N > 1: the batch is sharded (8192 rows per rank, weak scaling); item
embeddings are all-gathered so every rank scores its users against the global batch of negatives, dV is
reduce-scattered back and dense gradients are all-reduced (two_tower_models_b200/distributed.py).

`value`  : pairs/s with the step's inputs already resident in HBM (a ring of batches larger than L2).
`e2e`    : pairs/s through the public module API from pinned HOST buffers: H2D of the 7 input tensors and
           a D2H read of the loss inside the timed region, every step.
`--impl reference` : the reference's CPU path (the oracle port of it, oracle/two_tower_oracle.py, the
           reference itself is not present on the GPU box) on all host cores, same config and metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HASH = 100_000
RING = 32  # distinct input batches cycled through: 32 x ~8.5 MB = 272 MB > 126 MB L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    # defaults: ~0.3 s of timed device work per leg, long enough for the 50 ms nvidia-smi clock sampler to see the load
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--d", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8192, help="rows per GPU")
    ap.add_argument("--features", type=int, default=128)
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="only the configs[1] line (skip the d=256 / history / MIPS legs)")
    ap.add_argument("--peer-ce", action="store_true",
                    help="N > 1: scoring kernels read the item shards from peer memory (no NCCL all-gather)")
    ap.add_argument("--workload", default="base", choices=["base", "history", "mips"],
                    help="base = BASELINE configs[1] (the driver's bench line); history = configs[2]; mips = configs[3]")
    ap.add_argument("--queries", type=int, default=65536)
    ap.add_argument("--corpus", type=int, default=1_000_000)
    ap.add_argument("--topk", type=int, default=100)
    ap.add_argument("--hist-len", type=int, default=50)
    ap.add_argument("--layers", type=int, default=2)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def make_batch(B, F, gen):
    return dict(
        user_id=torch.randint(0, HASH, (B,), generator=gen),
        user_features=torch.randn(B, F, generator=gen),
        user_history=torch.randint(0, HASH, (B, 8), generator=gen),
        item_id=torch.randint(0, HASH, (B,), generator=gen),
        item_features=torch.randn(B, F, generator=gen),
        position=torch.randint(0, 100, (B,), generator=gen),
        labels=torch.randint(0, 2, (B, 1), generator=gen).float(),
    )


ORDER = ["user_id", "user_features", "user_history", "item_id", "item_features", "position", "labels"]


def batch_bytes(b):
    return sum(t.numel() * t.element_size() for t in b.values())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        # under load = samples in the upper half (the sampler also sees idle gaps around the region)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def load_reference():
    """The UNMODIFIED reference, installed with `pip install --target baseline/_ref /root/reference` (package `src`);
    None when it is not present (then the oracle port stands in, kind = "port")."""
    p = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(p, "src")):
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    import importlib

    try:
        return (importlib.import_module("src.two_tower_base_retrieval").TwoTowerBaseRetrieval,
                importlib.import_module("src.baseline_mips_module").BaselineMIPSModule)
    except Exception:
        return None


def reference_step_fn(d, F, B):
    """(callable(batch) -> loss, kind, description): one train_forward + backward of the reference's CPU path."""
    ref = load_reference()
    if ref is not None:
        Model, Mips = ref
        torch.manual_seed(0)
        m = Model(100, HASH, d, F, HASH, d, F, [1.0], Mips(16, d))

        def step(b):
            m.zero_grad(set_to_none=True)
            loss = m.train_forward(*[b[k] for k in ORDER])
            loss.backward()
            return float(loss)

        return step, "reference", "the unmodified reference (baseline/_ref, src.two_tower_base_retrieval) on CPU, fp32, autograd backward"
    import oracle

    params, uvw = oracle_params(d, F)

    def step(b):
        loss, _ = oracle.base_train_forward_with_grads(params, uvw, b)
        return float(loss)

    return step, "port", "oracle port of the reference CPU path (baseline/_ref absent), fp32, autograd backward"


def run_reference(args, rank, world):
    """The reference's own CPU path on all host threads, same config / metric.  N > 1: rank 0 alone runs it."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    d, F = args.d, args.features
    # same config as the GPU arm: ONE process on the concatenated global batch (gpus x batch rows, every user scored against
    # all of them).  The reference materialises ~5 [B, B] fp32 matrices (scores, log-softmax, their gradients); if the host
    # cannot hold them the largest multiple of the per-GPU batch that fits is timed instead and the line says so.
    want = args.batch * max(1, args.gpus)
    B = want
    try:
        import psutil

        avail = psutil.virtual_memory().available
        while B > args.batch and 5.5 * 4.0 * B * B > 0.8 * avail:
            B -= args.batch
    except Exception:
        pass
    step, kind, what = reference_step_fn(d, F, B)
    if B != want:
        what += f"; host memory holds a global batch of {B} rows, not the {want} of the GPU arm (cost per pair grows with the batch: this OVERSTATES the reference's pairs/s)"
    gen = torch.Generator().manual_seed(1)
    batches = [make_batch(B, F, gen) for _ in range(2)]
    W = max(1, min(args.warmup, 2)) if B <= 16384 else 1
    for i in range(W):
        step(batches[i % 2])
    steps = max(1, min(args.steps, 5 if B <= 16384 else 2))
    t0 = time.perf_counter()
    for i in range(steps):
        step(batches[i % 2])
    dt = (time.perf_counter() - t0) / steps
    val = B / dt
    cfg = workload_config(args, max(1, args.gpus))
    cfg["global_batch"] = B
    cfg["reference_process"] = "single process, whole global batch"
    print(json.dumps({
        "impl": "reference", "metric": "user-item pairs/sec through train_forward (fwd+bwd)", "value": val,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": W,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{steps} full steps of B={B}: {what}"},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def workload_config(args, world, d=None):
    d = args.d if d is None else d
    return {
        "workload": f"TwoTowerBaseRetrieval.train_forward+backward d={d} F={args.features} batch={args.batch}/GPU "
                    f"hash={HASH} T=1 (BASELINE configs[{4 if d == 256 else 1}])",
        "global_batch": args.batch * world, "d": d, "parallelism": f"dp{world}" if world > 1 else "single",
        "negatives": ("in-batch, read in place from peer memory over NVLink" if getattr(args, "peer_ce", False)
                      else "in-batch, all-gathered over NCCL") if world > 1 else "in-batch",
        "l2": f"input ring of {RING} batches > 126 MB L2",
        "step": "train_forward + loss.backward(), gradients of all 14 parameter tensors, weights re-cast to bf16 every step",
    }


def oracle_params(d, F):
    """Reference-default-initialised parameters (seed 0) as a state_dict; built from torch.nn hosts."""
    import two_tower_models_b200 as tt

    torch.manual_seed(0)
    m = tt.TwoTowerBaseRetrieval(100, HASH, d, F, HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d))
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, torch.tensor([1.0])


def run_mips(args, rank, world, local_rank):
    """BASELINE configs[3]: BaselineMIPSModule, 1M x 128 corpus, 65 536 queries, top-100 (queries/s).
    N > 1: queries are sharded over replicas of the corpus, no collective.  Returns the result dict on rank 0."""
    import torch.distributed as dist
    import two_tower_models_b200 as tt
    from two_tower_models_b200 import ops

    dev = torch.device("cuda", local_rank)
    Q, C, d, k = args.queries, args.corpus, 128, args.topk
    K, W = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    torch.manual_seed(0)
    mips = tt.BaselineMIPSModule(C, d).to(dev)
    gen = torch.Generator().manual_seed(1 + rank)
    q_host = torch.randn(Q, d, generator=gen).pin_memory()
    q_dev = q_host.to(dev)
    c16 = mips._packed.get("corpus", mips.corpus)

    def step(q):
        return ops.mips_topk(q, mips.corpus, c16, k)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step(q_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.profile_report()
    ops.profile_kernels(True)
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        idx, sc = step(q_dev)
    e1.record()
    barrier()
    launches = ops.launch_count() - l0
    spans = ops.profile_report()
    ops.profile_kernels(False)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = Q * world / (ms_step * 1e-3)
    # e2e: queries from pinned host memory, indices + scores back to the host, through the module API
    idx_host = torch.empty((Q, k), dtype=torch.int64).pin_memory()
    sc_host = torch.empty((Q, k), dtype=torch.float32).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        i2, s2, _ = mips(q_host.to(dev, non_blocking=True), k)
        idx_host.copy_(i2, non_blocking=True)
        sc_host.copy_(s2, non_blocking=True)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    tt2 = torch.tensor([e2e_ms], device=dev)
    if world > 1:
        dist.all_reduce(tt2, op=dist.ReduceOp.MAX)
    e2e_ms = float(tt2.item())
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    pk, pk_src = peaks()
    flops = 2.0 * Q * C * d
    screen_ms = spans.get("mips_screen_kernel", (ms_step * K, K))
    screen_ms = screen_ms[0] / max(screen_ms[1], 1)
    tf = flops / (screen_ms * 1e-3) / 1e12
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])  # a 40 ms kernel under the power cap: sustained peak
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count())
        cq = q_host[:1024].clone()
        cc = mips.corpus.cpu()
        ref = load_reference()
        if ref is not None:  # the unmodified reference module with the same corpus
            rm = ref[1](C, d)
            rm.corpus = cc
            oracle_topk, kind = (lambda: rm(cq, k)), "reference"
        else:
            oracle_topk, kind = (lambda: torch.topk(cq @ cc.t(), k, dim=1)), "port"  # the reference's own two calls (:57-61)
        oracle_topk()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            oracle_topk()
        dt = (time.perf_counter() - t0) / n
        cpu = {"value": 1024 / dt, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{n} x one 1024-query chunk against the full {C}-row corpus (matmul + topk + gather, fp32; the full "
                         f"[{Q},{C}] score matrix does not fit host memory)"}
    # parity on a 1-in-64 sample of the queries: recall of the reference's top-k (fp32 scores on the host)
    parity = None
    if world == 1 and not args.no_cpu_baseline:
        sel = torch.arange(0, Q, 64)[:256]
        sc_ref, idx_ref = torch.topk(q_host[sel] @ mips.corpus.cpu().t(), k, dim=1)
        got = idx[sel.to(dev)].cpu()
        hit = sum(len(set(got[i].tolist()) & set(idx_ref[i].tolist())) for i in range(len(sel)))
        parity = {"recall_vs_fp32_topk": hit / float(len(sel) * k), "queries_checked": int(len(sel)),
                  "score_rel_err_max": float(((sc[sel.to(dev)].cpu() - sc_ref).abs().max() / sc_ref.abs().max()))}
    return ({
        "metric": "MIPS queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": f"BaselineMIPSModule corpus={C} d={d} queries={Q}/GPU top-{k} (BASELINE configs[3])",
                   "screening": "bf16 tensor-core scores, k+32 candidates re-scored in fp32",
                   "l2": "corpus copies (256 MB bf16 + 512 MB fp32) exceed the 126 MB L2"},
        "e2e": {"value": Q * world / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": Q * d * 4,
                "d2h_bytes_per_step": Q * k * 12, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "mips_screen_kernel", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": tf / peak_tf, "traffic": None, "peak_source": pk_src, "flops_per_launch": flops,
                     "kernels": {k_: {"ms": v[0] / max(v[1], 1)} for k_, v in spans.items()}},
        "cpu_baseline": cpu, "parity": parity, "clocks": clocks,
    })


def run_history(args, rank, world, local_rank):
    """BASELINE configs[2]: TwoTowerWithUserHistoryEncoder, hist_len 50, 2 attention layers, d=128, B=8192."""
    import two_tower_models_b200 as tt
    from two_tower_models_b200 import ops
    from two_tower_models_b200.graph import GraphedTrainStep

    assert world == 1, "the history workload is benchmarked on one GPU"
    dev = torch.device("cuda", local_rank)
    B, d, F, H, L = args.batch, 128, args.features, args.hist_len, args.layers
    K, W = max(3, min(args.steps, 50)), max(min(args.warmup, 10), 3)
    torch.manual_seed(0)
    model = tt.TwoTowerWithUserHistoryEncoder(100, HASH, d, F, H, HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d),
                                              num_attention_heads=4, num_attention_layers=L).to(dev)
    gen = torch.Generator().manual_seed(1)
    ring = []
    for _ in range(8):
        b = make_batch(B, F, gen)
        b["user_history"] = torch.randint(0, HASH, (B, H), generator=gen)
        ring.append({k: v.pin_memory() for k, v in b.items()})
    dring = [{k: v.to(dev) for k, v in b.items()} for b in ring]
    gstep = GraphedTrainStep(model, dring[0])
    dring = [gstep.new_slot(b) for b in dring]  # packed like the captured inputs: one copy per step
    ring = [gstep.new_slot(b, device="cpu", pin_memory=True) for b in ring]
    for i in range(W):
        gstep(dring[i % 8])
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        loss = gstep(dring[(W + i) % 8])
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / K
    loss_host = torch.zeros(K, dtype=torch.float32).pin_memory()
    t0 = time.perf_counter()
    for i in range(K):
        gstep.load(ring[i % 8], non_blocking=True)
        loss_host[i].copy_(gstep.replay(), non_blocking=True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    clocks = sampler.stop()
    ops.profile_report()
    ops.profile_kernels(True)
    model.zero_grad(set_to_none=True)
    l0 = ops.launch_count()
    lz = model.train_forward(*[dring[0][k] for k in ORDER])
    lz.backward()
    launches = ops.launch_count() - l0
    spans = ops.profile_report()
    ops.profile_kernels(False)
    cpu = None
    if not args.no_cpu_baseline:
        import oracle

        torch.set_num_threads(os.cpu_count())
        params = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
        pe = model.user_history_encoder.positional_embeddings.cpu()
        hb = {k: ring[0][k][:1024].clone() for k in ORDER}
        oracle.history_train_forward_with_grads(params, torch.tensor([1.0]), hb, 4, pe)
        t0 = time.perf_counter()
        oracle.history_train_forward_with_grads(params, torch.tensor([1.0]), hb, 4, pe)
        dt = time.perf_counter() - t0
        cpu = {"value": 1024 / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "one fwd+bwd step of a 1024-row slice of the batch (oracle port, fp32); the B x B loss part "
                         "is 64x smaller than at B=8192, the encoder part scales linearly"}
    enc_flops = L * (2.0 * B * H * d * 3 * d + 4.0 * B * H * H * d + 2.0 * B * H * d * d)
    # roofline of the whole step: algorithmic flops (no recompute, full H x H attention of every layer, backward
    # = 2 x forward) over the step time, against the measured bf16 peak
    base_flops = 2.0 * B * B * d + 2.0 * B * (2 * F * 256 + 2 * 256 * d + (2 * d + 2 * d) * d + 2 * d * d)
    step_flops = 3.0 * (enc_flops + base_flops)
    pk, pk_src = peaks()
    peak_tf = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    tf = step_flops / (ms_step * 1e-3) / 1e12
    return ({
        "metric": "user-item pairs/sec through train_forward (fwd+bwd)", "value": B / (ms_step * 1e-3), "unit": "pairs/s",
        "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"TwoTowerWithUserHistoryEncoder.train_forward+backward d={d} F={F} batch={B} hist_len={H} "
                               f"layers={L} heads=4 (BASELINE configs[2])"},
        "e2e": {"value": B / (e2e_ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": batch_bytes({k: ring[0][k] for k in ORDER}),
                "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches) * K, "launches_per_step": int(launches), "cuda_graph": True,
        "kernel_ms_per_step": {k_: v[0] for k_, v in spans.items()},
        "encoder_algorithmic_gflop_fwd": enc_flops / 1e9, "cpu_baseline": cpu, "clocks": clocks, "loss": float(loss.item()),
        "roofline": {"bound": "tensor", "kernel": "whole step (encoder projections + attention + towers + B x B loss)",
                     "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf, "traffic": None,
                     "peak_source": pk_src, "flops_per_step": step_flops,
                     "note": "algorithmic flops of forward + backward (3 x forward) / step time"},
    })


def run_base(args, rank, world, local_rank, d, light):
    """One BASELINE configs[1]-shaped leg (TwoTowerBaseRetrieval, train_forward + backward) at embedding dim d; the result
    dict on rank 0, None elsewhere.  light = secondary leg (bounded steps; no e2e / CPU / optimizer parts)."""
    import torch.distributed as dist
    import two_tower_models_b200 as tt
    from two_tower_models_b200 import ops

    dev = torch.device("cuda", local_rank)
    B, F = args.batch, args.features
    K, W = args.steps, max(args.warmup, 3)
    if light:  # secondary leg: a bounded number of steps, no end-to-end / CPU / optimizer parts
        K = max(3, min(K, 50))

    torch.manual_seed(0)
    model = tt.TwoTowerBaseRetrieval(100, HASH, d, F, HASH, d, F, [1.0], tt.BaselineMIPSModule(16, d)).to(dev)
    if world > 1:
        from two_tower_models_b200 import distributed as ttd

        ttd.enable_data_parallel(model, peer_memory=True if args.peer_ce else None)

    gen = torch.Generator().manual_seed(1 + rank)
    host_ring = [{k: v.pin_memory() for k, v in make_batch(B, F, gen).items()} for _ in range(RING)]
    dev_ring = [{k: v.to(dev) for k, v in b.items()} for b in host_ring]
    in_bytes = batch_bytes({k: host_ring[0][k] for k in ORDER})

    sync = model._dp.sync_gradients if world > 1 else None
    use_graph = not args.no_graph
    gstep_holder = []
    if use_graph:
        from two_tower_models_b200.graph import GraphedTrainStep

        gstep = GraphedTrainStep(model, dev_ring[0], post_backward=sync)
        gstep_holder.append(gstep)
        # ring entries and the pinned host batches packed like the captured inputs: one copy per step instead of seven
        dev_ring = [gstep.new_slot(b) for b in dev_ring]
        host_ring = [gstep.new_slot(b, device="cpu", pin_memory=True) for b in host_ring]

        def step(b):  # D2D copy of the batch into the captured inputs + one graph launch
            return gstep(b)
    else:
        def step(b):
            model._packed.invalidate()
            model.zero_grad(set_to_none=True)
            loss = model.train_forward(*[b[k] for k in ORDER])
            loss.backward()
            if sync is not None:
                sync(model)
            return loss.detach()

    def eager_step(b):
        model._packed.invalidate()
        model.zero_grad(set_to_none=True)
        loss = model.train_forward(*[b[k] for k in ORDER])
        loss.backward()
        if sync is not None:
            sync(model)
        return loss.detach()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    for i in range(W):
        step(dev_ring[i % RING])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    eager_step(dev_ring[0])
    torch.cuda.synchronize()
    launches_per_step = ops.launch_count() - l0  # the captured graph replays exactly these launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        loss = step(dev_ring[(W + i) % RING])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * K
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / K
    value = B * world / (ms_step * 1e-3)
    loss_val = float(loss.item())

    # per-kernel device time: the library brackets each of its kernel launches with CUDA events on the
    # launching stream (eager launches of the very same kernels on the same inputs; events cannot be recorded
    # inside a replayed graph)
    ops.profile_report()
    ops.profile_kernels(True)
    n_prof = max(3, min(K, 10))
    for i in range(n_prof):
        # keep the device busy while the host enqueues the step, so that every start event is processed right before
        # its kernel (on an idle stream the event is stamped at once and the host's launch gap lands inside the span)
        torch.cuda._sleep(6_000_000)
        eager_step(dev_ring[(W + i) % RING])
        ops.profile_null_span()
    spans = ops.profile_report()
    ops.profile_kernels(False)
    # ... and inside the replayed graph: a second capture of the same step with an external event-record node on
    # either side of every kernel; each replay re-stamps the events.  These are the durations the kernels have in the
    # timed region (back to back, no launch gap inside the span); the eager spans above stay in the line beside them.
    span_src = "eager launches behind a busy device (CUDA events around every kernel)"
    eager_spans = None
    if use_graph and world == 1:
        try:
            ops.profile_kernels(2)
            def after_backward(m):  # an empty kernel at the end of the step: the overhead of a span's event pair
                if sync is not None:
                    sync(m)
                ops.profile_null_span()

            pstep = GraphedTrainStep(model, dev_ring[0], post_backward=after_backward)
            ops.profile_kernels(False)
            ops.profile_report()  # drop the eager warm-up spans of that capture
            acc = {}
            for i in range(n_prof):
                pstep(dev_ring[(W + i) % RING])
                for k_, (ms_, n_) in ops.profile_graph_report().items():
                    t_, c_ = acc.get(k_, (0.0, 0))
                    acc[k_] = (t_ + ms_, c_ + n_)
            ops.profile_graph_report(clear=True)
            del pstep
            if acc:
                eager_spans, spans = spans, acc
                span_src = "CUDA-graph replays of the step with event-record nodes around every kernel"
        except Exception as ex:  # noqa: BLE001 - keep the eager spans
            ops.profile_kernels(False)
            span_src += f" (graph spans unavailable: {ex})"

    # ---------------- end to end from pinned host buffers (`e2e`) ----------------
    loss_host = torch.zeros(K + W, dtype=torch.float32).pin_memory()
    e2e_ms = None
    if not light:  # double-buffered H2D on a copy stream; the step (graph replay) starts with a D2D into its inputs
        copy_stream = torch.cuda.Stream(device=dev)
        slots = ([gstep_holder[0].new_slot() for _ in range(2)] if use_graph else
                 [{k: torch.empty_like(v, device=dev) for k, v in host_ring[0].items()} for _ in range(2)])
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                if use_graph:  # one H2D copy of the packed batch
                    slots[s]["_flat"].copy_(host_ring[i % RING]["_flat"], non_blocking=True)
                else:
                    for k in ORDER:
                        slots[s][k].copy_(host_ring[i % RING][k], non_blocking=True)
                ready[s].record(copy_stream)

        def e2e_steps(n0, n):
            cur = torch.cuda.current_stream()
            upload(n0)
            for i in range(n0, n0 + n):
                if i + 1 < n0 + n:
                    upload(i + 1)  # prefetch the next batch while this one computes
                s = i % 2
                cur.wait_event(ready[s])
                l = step(slots[s])
                freed[s].record(cur)
                loss_host[i].copy_(l.detach(), non_blocking=True)  # D2H read of the step's result

        for s in range(2):
            freed[s].record(torch.cuda.current_stream())
        e2e_steps(0, W)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        e2e_steps(W, K)
        e1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[1].item()) / K  # host wall clock between the syncs: includes every copy and launch
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    del gstep_holder[:]
    if rank != 0:
        return None

    # ---------------- roofline of the dominant kernel (B x N scoring) ----------------
    pk, pk_src = peaks()
    N = B * world
    kern = {}
    alg = {k_: 2.0 * B * N * d for k_ in ("ce_fwd_kernel", "ce_bwd2_kernel_dU", "ce_bwd2_kernel_dV", "ce_bwd3_kernel_dU",
                                           "ce_bwd3_kernel_dV", "ce_bwd3x_kernel_dU", "ce_bwd3x_kernel_dV")}
    # an event pair around an EMPTY kernel measures a few microseconds: the span of every kernel carries that much on
    # top of its execution time, so the per-kernel rates use span - null span (both are in the line)
    null_tot, null_cnt = spans.pop("null_kernel", (0.0, 0))
    span_overhead = null_tot / null_cnt if null_cnt else 0.0
    if eager_spans:
        eager_spans.pop("null_kernel", None)
    for name, (tot, cnt) in spans.items():
        per = tot / max(cnt, 1)
        kern[name] = {"ms": max(per - span_overhead, 0.25 * per), "span_ms": per, "launches_per_step": cnt / n_prof}
        if name in alg:
            kern[name]["tflops"] = alg[name] / (kern[name]["ms"] * 1e-3) / 1e12
            kern[name]["tflops_raw_span"] = alg[name] / (per * 1e-3) / 1e12
    scoring = [k for k in kern if k in alg]
    dom = max(scoring, key=lambda k: kern[k]["ms"]) if scoring else None
    # the scoring kernels are timed one by one behind a busy device (tens of microseconds each): the BURST peak applies
    peak_tf = pk["bf16_tflops"]
    roofline = None
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch from the committed `ncu --set full` capture of this very workload
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        if world == 1 and tj.get("config") == f"base B={B} d={d} F={F} n_gpus=1" and dom in tj["bytes_per_launch"]:
            traffic, traffic_src = tj["bytes_per_launch"][dom], f"profiles/ncu_traffic.json ({tj['source']})"
    except Exception:
        pass
    if dom:
        sc_ms = sum(kern[k]["ms"] for k in scoring)
        roofline = {
            "bound": "tensor", "kernel": dom, "achieved": kern[dom]["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
            "frac": kern[dom]["tflops"] / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": 2.0 * (B + N) * d + 4.0 * N * d,
            "peak_source": f"{pk_src} (burst cuBLAS bf16 peak: the kernel is timed alone, tens of microseconds)",
            "frac_of_sustained_peak": kern[dom]["tflops"] / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
            "flops_per_launch": alg[dom],
            "note": "algorithmic flops 2*B*N*d per launch (dS . Y only; the recomputed S = X Y^T is not counted)",
            "step": {"ms": ms_step, "algorithmic_gflop": (6.0 * B * N * d + 6.0 * B * (2 * F * 256 + 2 * 256 * d + 4 * d * d)) / 1e9,
                     "tflops": (6.0 * B * N * d + 6.0 * B * (2 * F * 256 + 2 * 256 * d + 4 * d * d)) / (ms_step * 1e-3) / 1e12,
                     "scoring_roofline_frac": 6.0 * B * N * d / (ms_step * 1e-3) / 1e12 / pk.get("bf16_tflops_sustained", pk["bf16_tflops"])},
            "scoring_all": {"ms": sc_ms, "tflops": 6.0 * B * N * d / (sc_ms * 1e-3) / 1e12},
            "kernels": kern,
            "kernel_timing": span_src + "; ms = span - span of an empty kernel measured the same way",
            "event_span_overhead_ms": span_overhead,
            "frac_raw_span": kern[dom]["tflops_raw_span"] / peak_tf,
            "kernels_eager_ms": ({k_: v_[0] / max(v_[1], 1) for k_, v_ in eager_spans.items()} if eager_spans else None),
            "device_ms_all_kernels": sum(v["ms"] * v["launches_per_step"] for v in kern.values()),
        }

    cpu, parity = None, None
    if world == 1 and not args.no_cpu_baseline and not light:
        torch.set_num_threads(os.cpu_count())
        ref_step, kind, what = reference_step_fn(d, F, B)
        hb = {k: host_ring[0][k].clone() for k in ORDER}
        ref_step(hb)
        n = 3
        t0 = time.perf_counter()
        for _ in range(n):
            ref_loss = ref_step(hb)
        dt = (time.perf_counter() - t0) / n
        cpu = {"value": B / dt, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"{n} full steps of B={B} on the host: {what}", "ms_per_step": dt * 1e3}
        # the same batch through the CUDA path: the loss has to agree with the fp32 CPU value (tolerance 1e-3)
        gpu_loss = float(eager_step(dev_ring[0]).item())
        parity = {"loss_gpu": gpu_loss, "loss_cpu_fp32": ref_loss, "loss_rel_vs_oracle": abs(gpu_loss - ref_loss) / abs(ref_loss),
                  "tolerance": 1e-3, "checker": kind}

    # ---------------- adjacent step: fused Adam over all parameters, timed on its own ----------------
    optim_info = None
    if world == 1 and not light:
        try:
            opt = tt.FusedAdam(model.parameters(), lr=1e-3)
            opt.step()  # creates the state
            torch.cuda.synchronize()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                opt.step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            g_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_opt):
                opt.step()
            for _ in range(10):  # the device idled during the host-side baseline: bring the clocks back up
                g_opt.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                g_opt.replay()
            e1.record()
            torch.cuda.synchronize()
            oms = e0.elapsed_time(e1) / 20
            numel = sum(p.numel() for p in model.parameters())
            gbs = 28.0 * numel / (oms * 1e-3) / 1e9
            optim_info = {"kernel": "adam_kernel", "ms": oms, "elements": numel, "bytes": 28 * numel, "achieved_gbs": gbs,
                          "hbm_peak_gbs": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"],
                          "note": "torch.optim.Adam semantics on dense gradients (reference train/train.py:179); not "
                                  "part of the timed step (SURVEY 8d excludes the optimizer)"}
        except Exception as exc:  # pragma: no cover - reported, never fatal for the bench line
            optim_info = {"error": str(exc)[:200]}

    out = {
        "metric": "user-item pairs/sec through train_forward (fwd+bwd)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, world, d),
        "scored_pairs_per_s": B * (B * world) * world / (ms_step * 1e-3),
        "e2e": None if e2e_ms is None else {"value": B * world / (e2e_ms * 1e-3), "unit": "pairs/s",
                                            "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms},
        "gpu_launches": int(launches), "launches_per_step": launches_per_step, "cuda_graph": use_graph,
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "clocks": clocks, "loss": loss_val,
        "optimizer_step": optim_info,
    }
    return out




def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def finish(out):
        if rank == 0:
            print(json.dumps(out), flush=True)
        sys.stdout.flush()
        if world > 1:
            os._exit(0)  # no collective teardown

    if args.workload == "mips":
        return finish(run_mips(args, rank, world, local_rank))
    if args.workload == "history":
        return finish(run_history(args, rank, world, local_rank))

    # the driver's line: BASELINE configs[1] (d = 128) - and, attached under "configs", the other BASELINE configurations
    # as extra timed legs: configs[4] (d = 256, the same sharded step), and on one GPU configs[2] (history) / configs[3] (MIPS)
    out = run_base(args, rank, world, local_rank, args.d, light=False)
    extra = {}
    if not args.no_extra_legs:
        def guarded(name, fn):
            try:
                r = fn()
            except Exception as exc:  # a secondary leg never takes the headline line down
                r = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            if rank == 0:
                extra[name] = r

        if args.d != 256:
            guarded("config5_d256", lambda: run_base(args, rank, world, local_rank, 256, light=True))
        if world == 1:
            guarded("history", lambda: run_history(args, rank, world, local_rank))
            guarded("mips", lambda: run_mips(args, rank, world, local_rank))
    if rank == 0:
        out["configs"] = extra
    finish(out)


if __name__ == "__main__":
    main()
