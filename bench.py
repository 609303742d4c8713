#!/usr/bin/env python
"""bench.py — headline benchmark of the MIPS hot path (BASELINE.json: "MIPS queries/s
(10M x 768, k=8) at 1/2/4/8 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one search of a batch of 1024 synthetic queries against the 10M x 768 bf16 memory
bank (row-sharded over the N ranks: strong scaling, the bank is the fixed job), k=8: query
prep + K1 (tcgen05 search) + local merge (+ one NCCL all-gather + final merge when N>1), all of it
ONE C-ABI call (`mips_search_sharded`) captured in ONE CUDA graph.
`value` times graph replays with the bank AND the queries resident in HBM; `e2e` times the same step
with the queries in pinned host memory and (D, I) read back to the host every step. `roofline` times
the K1 launches of the same step with CUDA events on its stream (an eager loop of the same K steps:
events cannot be read back from inside a graph). `parity` checks the timed path against a chunked
torch fp32 brute force over the same bf16-rounded rows at every N. Prints ONE JSON line (rank 0).

`--impl reference` times the reference's CPU search path (flat exact search: sgemm + top-k + merge
over 1M-row chunks, what faiss-cpu IndexFlatIP does; restated in oracle/mips_oracle.py) on the host
cores over the FULL 10M-row bank, every reported step a real step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ROWS, DIM, NQ, TOPK = 10_000_000, 768, 1024, 8
METRIC = "mips_queries_per_s"
UNIT = "queries/s"


def workload_name(n: int, d: int, nq: int, k: int) -> str:
    return (f"{n}x{d} bf16 memory bank, batch {nq} queries, k={k}, exact inner-product search "
            f"(BASELINE config 3)")


def _peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2])), power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": float(max(power)), "power_w_median": float(np.median(power)), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arm
def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_sample_leg(nq: int, d: int, k: int, n_full: int, budget_s: float = 15.0) -> dict:
    """`cpu_baseline` of our own arm: the reference's CPU search path (oracle port, all host threads) on a
    BOUNDED sample — all nq queries against one 1M-row chunk of the bank, repeated for ~budget_s — scaled to
    the full bank (the full-bank measurement is what `--impl reference` does)."""
    import torch
    from oracle import mips_oracle as o   # test infrastructure, used here only as the CPU baseline

    cores = _host_threads()
    torch.set_num_threads(cores)
    rows = min(1_000_000, n_full)
    g = torch.Generator().manual_seed(99)
    xb = torch.randn((rows, d), generator=g).bfloat16().float().numpy()
    xq = torch.randn((nq, d), generator=g).bfloat16().float().numpy()
    o.flat_search_chunked(xb[:10000], xq[:8], k, chunk_rows=10000)      # warm the thread pool
    t0 = time.perf_counter()
    reps = 0
    while True:
        o.flat_search_chunked(xb, xq, k, chunk_rows=rows)
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 5:
            break
    t_chunk = (time.perf_counter() - t0) / reps
    qps = nq / (t_chunk * n_full / rows)
    return {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{nq} queries x {rows} of {n_full} rows x {d} (fp32 values of the bf16 bank), k={k}: torch sgemm + "
                      f"top-k per chunk ({t_chunk:.3f} s per chunk, {reps} reps), time scaled by {n_full / rows:.0f}x "
                      f"to the full bank; `--impl reference` measures the full bank"}


def run_reference(args) -> None:
    """The reference arm: every step is a REAL exact search over the FULL bank on the host cores — matmul +
    top-k per 1M-row chunk + running merge, i.e. what the flat faiss-cpu index behind `Mips.search`
    (sotasum/mips.py:383-386) computes and what sotasum/retriever_lightning.py:304-305 spells in torch. The
    queries per step are a bounded sample (a power-of-two share of the 1024, chosen by a one-chunk calibration so
    that K steps + W warm-ups end within ~3 minutes); nothing is extrapolated: value = queries of a step / its
    measured time."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import mips_oracle as o   # the CPU restatement of the reference path IS what this arm times

    t_wall0 = time.perf_counter()
    cores = _host_threads()
    torch.set_num_threads(cores)
    n, d, nq, k = args.rows, DIM, args.nq, args.k
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    chunk = min(1_000_000, n)
    n_chunks = (n + chunk - 1) // chunk
    # the bank: one chunk of N(0,1) rows (bf16 values up-cast, like the GPU arm's bank); the other chunks are
    # column rotations of it (distinct rows, same statistics; sgemm / top-k time does not depend on the values)
    avail_gb = 0.0
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable"):
                avail_gb = int(line.split()[1]) / 1e6
    except Exception:
        pass
    full_resident = avail_gb > n * d * 4 / 1e9 + 8.0
    g = torch.Generator().manual_seed(1234)
    base = torch.randn((chunk, d), generator=g).bfloat16().float()
    xq_all = torch.randn((nq, d), generator=torch.Generator().manual_seed(4321)).bfloat16().float().numpy()
    if full_resident:
        bank = torch.empty((n, d), dtype=torch.float32)
        for c in range(n_chunks):
            m = min(chunk, n - c * chunk)
            bank[c * chunk:c * chunk + m] = torch.roll(base, shifts=7 * c, dims=1)[:m]
        bank_np = bank.numpy()
        bank_note = f"full {n}x{d} fp32 bank resident in host memory ({n * d * 4 / 1e9:.1f} GB)"
    else:
        bank_np = None
        bank_note = (f"host memory too small for the full fp32 bank ({avail_gb:.0f} GB available): the same 1M-row chunk "
                     f"is searched {n_chunks} times per step with shifted ids")
    # calibration: all queries against one chunk
    o.flat_search_chunked(base.numpy()[:10000], xq_all[:8], k, chunk_rows=10000)
    t0 = time.perf_counter()
    o.flat_search_chunked(base.numpy(), xq_all, k, chunk_rows=chunk)
    t_chunk = time.perf_counter() - t0
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", "170"))
    nq_s = nq
    while nq_s > 32 and t_chunk * n_chunks * (nq_s / nq) * (steps + warmup) > budget_s:
        nq_s //= 2
    xq = np.ascontiguousarray(xq_all[:nq_s])

    def step():
        if bank_np is not None:
            return o.flat_search_chunked(bank_np, xq, k, chunk_rows=chunk)
        best = None
        for c in range(n_chunks):
            D, I = o.flat_search_chunked(base.numpy(), xq, k, chunk_rows=chunk)
            best = (D, I + c * chunk) if best is None else o.merge_topk_pair(best, (D, I + c * chunk), k)
        return best

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        D, I = step()
    t_step = (time.perf_counter() - t0) / steps
    assert I.shape == (nq_s, k) and (np.diff(D, axis=1) <= 0).all()
    qps = nq_s / t_step
    leg = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"every step searches {nq_s} of the {nq} queries against the FULL bank ({bank_note}), k={k}: torch "
                     f"sgemm + top-k per 1M-row chunk + running merge (oracle/mips_oracle.py flat_search_chunked, the "
                     f"flat exact search behind sotasum/mips.py:383-386); {steps} timed + {warmup} warm-up steps, "
                     f"{t_step:.2f} s per step, nothing extrapolated"}
    # the reference's own first-party exact path (inner_product, sotasum/mips.py:552-560: sgemm + FULL argsort) on
    # BASELINE config 1 (100k x 768, 32 queries): the reference's code itself when /root/reference is here
    c1 = None
    try:
        xb1, xq1 = base.numpy()[:100_000], xq_all[:32]
        fn, kind = o.inner_product, "port"
        ref_src = Path("/root/reference/sotasum/mips.py")
        if ref_src.exists():
            import ast
            tree = ast.parse(ref_src.read_text())
            node = next(x for x in tree.body if isinstance(x, ast.FunctionDef) and x.name == "inner_product")
            ns = {"np": np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), str(ref_src), "exec"), ns)
            fn, kind = ns["inner_product"], "reference"
        t0 = time.perf_counter()
        fn(xq1.copy(), xb1.copy(), 8, normalize=False)
        c1 = {"workload": "BASELINE config 1: inner_product (mips.py:552-560), 100000x768 fp32, 32 queries, k=8",
              "kind": kind, "queries_per_s": 32 / (time.perf_counter() - t0)}
    except Exception as e:  # informative leg only
        c1 = {"error": repr(e)}
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n, d, nq, k), "queries_per_step": nq_s, "host_threads": cores},
            "cpu_baseline": leg, "c1_inner_product": c1,
            "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_wall0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- parity
def parity_check(torch, dist, m, idx, sh, graph, rows, n, d, k, world, rank, dev) -> dict:
    """Driver-visible parity of the TIMED path at this N: 256 queries (128 planted = bank rows spread over
    all shards + noise, 128 of the bench's random queries) through the same captured step, against a chunked
    torch fp32 brute force over the SAME bf16-rounded rows (each rank over its shard; per-rank top-k
    all-gathered and merged by (score desc, id asc)). ids must agree except inside ties (score gap <= 1e-4
    relative, the check_topk rule of the oracle); all ranks must hold bit-identical (D, I)."""
    nq_p = graph.nq
    n_pl = min(128, nq_p // 2)
    gen = torch.Generator(device=dev).manual_seed(777 + rank)
    # planted rows: n_pl / world from every shard
    per = (n_pl + world - 1) // world
    loc = torch.randint(0, len(rows), (per,), generator=gen, device=dev)
    planted = idx.reconstruct_n(0, 1, as_torch=True).new_empty((per, d))
    for j, r in enumerate(loc.tolist()):
        planted[j] = idx.reconstruct_n(int(r), 1, as_torch=True)[0]
    planted_ids = loc + rows.start
    if world > 1:
        allp = torch.empty((world * per, d), dtype=torch.float32, device=dev)
        alli = torch.empty((world * per,), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allp, planted.contiguous())
        dist.all_gather_into_tensor(alli, planted_ids.contiguous())
    else:
        allp, alli = planted, planted_ids
    allp, alli = allp[:n_pl], alli[:n_pl]
    qg = torch.Generator(device="cpu").manual_seed(4321)
    rnd = torch.randn((nq_p - n_pl, d), generator=qg, dtype=torch.float32).to(dev)
    noise = torch.randn((n_pl, d), generator=torch.Generator(device="cpu").manual_seed(99), dtype=torch.float32).to(dev)
    Q = torch.cat([allp + 0.05 * noise, rnd]).bfloat16().float()        # the search rounds queries to bf16: same inputs
    out = graph.replay(Q)
    torch.cuda.synchronize()
    D, I = out["scores"].clone(), out["ids"].clone()
    # brute force over this rank's shard
    best_s = torch.full((nq_p, k), -float("inf"), device=dev)
    best_i = torch.full((nq_p, k), -1, dtype=torch.int64, device=dev)
    step = 500_000
    for s0 in range(0, len(rows), step):
        m_rows = min(step, len(rows) - s0)
        X = idx.reconstruct_n(s0, m_rows, as_torch=True)
        S = Q @ X.T
        v, i = S.topk(min(k, m_rows), dim=1)
        cs = torch.cat([best_s, v], 1)
        ci = torch.cat([best_i, i + (rows.start + s0)], 1)
        o_ = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = cs.gather(1, o_), ci.gather(1, o_)
        del X, S
    if world > 1:
        gs = torch.empty((world * nq_p, k), device=dev)
        gi = torch.empty((world * nq_p, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gs, best_s.contiguous())
        dist.all_gather_into_tensor(gi, best_i.contiguous())
        cs = gs.view(world, nq_p, k).permute(1, 0, 2).reshape(nq_p, -1)
        ci = gi.view(world, nq_p, k).permute(1, 0, 2).reshape(nq_p, -1)
        o_ = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]
        best_s, best_i = cs.gather(1, o_), ci.gather(1, o_)
        allD = torch.empty((world * nq_p, k), device=dev)
        allI = torch.empty((world * nq_p, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allD, D.contiguous())
        dist.all_gather_into_tensor(allI, I.contiguous())
        ranks_identical = bool((allD.view(world, -1) == allD.view(world, -1)[0]).all() and
                               (allI.view(world, -1) == allI.view(world, -1)[0]).all())
    else:
        ranks_identical = True
    scale = best_s.abs().clamp_min(1.0)
    diff = I != best_i
    tie_ok = (D - best_s).abs() <= 1e-4 * scale          # a different id is a tie iff the scores agree
    hard = int((diff & ~tie_ok).sum())
    # id SETS must agree outside ties too (a swap of two tied ids changes positions, not the set)
    score_err = float(((D - best_s).abs() / scale).max())
    planted_found = int((I[:n_pl, 0] == alli).sum())
    ok = hard == 0 and score_err <= 1e-4 and planted_found == n_pl and ranks_identical
    return {"ok": bool(ok), "queries": nq_p, "planted": n_pl, "planted_found_rank1": planted_found,
            "id_mismatches": int(diff.sum()), "id_mismatches_outside_ties": hard, "max_score_rel_err": score_err,
            "ranks_identical": ranks_identical,
            "against": "chunked torch fp32 matmul+topk over the same bf16-rounded rows, per-shard top-k merged"}


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    import retrieval_augmented_mds_b200 as m
    from retrieval_augmented_mds_b200 import _lib
    from retrieval_augmented_mds_b200.sharded import ShardedFlatIndex, balanced_range

    import faulthandler
    # a stuck collective must end the run (with the stacks of every thread on stderr), not hold the box
    faulthandler.dump_traceback_later(int(os.environ.get("BENCH_WATCHDOG_S", "600")), exit=True)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the search path is CUDA only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    json_fd = 1
    if world > 1:
        # NCCL's communicator lines must stay visible to whoever launched the run (its NCCL_DEBUG level is
        # respected; INFO / INIT when none is set) and stdout must stay the ONE JSON line: everything this
        # process writes to fd 1 (NCCL logs to the C stdout) is routed to stderr, the JSON line goes to the
        # original stdout
        if "NCCL_DEBUG" not in os.environ:
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    if args.gpus != world and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)

    n, d, nq, k = args.rows, DIM, args.nq, args.k
    rows = balanced_range(n, rank, world)
    idx = m.B200FlatIndex(d, m.METRIC_INNER_PRODUCT, dtype="bf16", device=dev, capacity=len(rows),
                          id_offset=rows.start)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    t_build0 = time.perf_counter()
    chunk = 500_000
    for s in range(0, len(rows), chunk):
        blk = torch.randn((min(chunk, len(rows) - s), d), generator=gen, device=dev, dtype=torch.float32)
        idx.add(blk)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t_build0
    qgen = torch.Generator(device="cpu").manual_seed(4321)       # identical queries on every rank
    xq_host = torch.randn((nq, d), generator=qgen, dtype=torch.float32).pin_memory()
    xq_dev = xq_host.to(dev)
    sh = None
    if world > 1:
        sh = ShardedFlatIndex(idx, exchange=args.exchange)
        cnt = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
        allc = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, cnt)
        sh.counts = [int(c) for c in allc.cpu().tolist()]

    # the timed step: ONE CUDA graph = query prep + K1 + local merge (+ ncclAllGather + final merge)
    use_graph = not args.no_graph and (sh is None or sh.exchange == "native")
    graph_host = None
    if use_graph:
        graph = sh.capture(nq, k, algo=args.algo) if sh is not None else idx.capture(nq, k, algo=args.algo)
        graph.xq.copy_(xq_dev)
        if sh is not None and args.e2e_graph_io:
            # e2e variant: the H2D copy of the queries (pinned host memory) and the D2H copies of (D, I) inside the
            # graph — one launch per step. Measured at N=2: 6.71 ms per step against 6.53 ms for copies issued
            # around the graph (memcpy nodes launch slower than the copy engine calls they replace): off by default
            graph_host = sh.capture(nq, k, algo=args.algo, host_io=True)
            graph_host.xq_host.copy_(xq_host)

    def step_eager():
        if sh is not None:
            return sh.search(xq_dev, k, algo=args.algo)
        return idx.search_ex(xq_dev, k, algo=args.algo)

    def step_device():
        if use_graph:
            graph.graph.replay()
            return graph.out
        return step_eager()

    D_pin = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    I_pin = torch.empty((nq, k), dtype=torch.int64).pin_memory()

    def step_e2e():
        if sh is None:
            idx.search_host(xq_host.numpy(), k, D=D_pin.numpy(), I=I_pin.numpy())   # one C-ABI call, host in / out
            return
        if graph_host is not None:
            graph_host.replay_host()             # H2D + step + D2H in one graph, then synchronize: (D, I) are in the
            return                               # pinned host tensors graph_host.out_host
        if use_graph:
            r = graph.replay(xq_host)            # H2D from pinned memory on the step's stream + the graph
        else:
            r = sh.search(xq_host.to(dev, non_blocking=True), k)
        D_pin.copy_(r["scores"], non_blocking=True)
        I_pin.copy_(r["ids"], non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    # clocks / throttle reasons are sampled under load from the warm-up to the end of the e2e loop: at
    # N=8 the K timed steps alone can be shorter than one nvidia-smi sampling period
    L = _lib.lib()
    l0 = L.mips_launch_count()
    step_eager()                                  # (also counts the kernels one step launches)
    torch.cuda.synchronize()
    launches_per_step = L.mips_launch_count() - l0
    # K1 duration (roofline) INSIDE the timed region: with profiling on, every capture of the step records its own
    # pair of CUDA events around the K1 launch on the step's stream (external event-record nodes of the graph), so
    # the timed loop replays K graphs of the same step — one per timed step — and leaves K K1 durations behind
    timed_graphs = None
    idx.set_profiling(True)
    if use_graph:
        try:
            timed_graphs = graph.add_replicas(args.steps)[1:]
        except Exception as e:                    # pragma: no cover - fall back to the eager profiling loop below
            print(f"[bench] K1 events inside the graph unavailable ({e!r}); timing K1 in an eager loop", file=sys.stderr)
            timed_graphs = None
            idx.set_profiling(True)               # (restart the slot counter)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    barrier()
    if timed_graphs is None and use_graph:
        idx.set_profiling(False)
    elif not use_graph:
        idx.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if timed_graphs is not None:
        for g in timed_graphs:
            g.replay()
        out = graph.out
    else:
        for _ in range(args.steps):
            out = step_device()
    e1.record()
    torch.cuda.synchronize()
    ms_total = e0.elapsed_time(e1)
    k1_ms, k1_n, k1_where = -1.0, 0, ""
    if timed_graphs is not None or not use_graph:
        k1_ms, k1_n = idx.k1_ms_total()
        k1_where = ("CUDA events recorded around every K1 launch of the timed region itself (external event nodes of "
                    "the K replayed graphs)" if use_graph else "CUDA events around every K1 launch of the timed region")
    idx.set_profiling(False)
    barrier()
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    launches = launches_per_step * args.steps     # a graph replay launches the kernels its capture enqueued
    final_ids = out["ids"].clone()
    final_scores = out["scores"].clone()
    eager_ms_step = None
    if k1_n == 0:
        # fallback: the same K steps, eager, CUDA events around each K1 launch on its stream
        barrier()
        idx.set_profiling(True)
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(args.steps):
            step_eager()
        ee1.record()
        torch.cuda.synchronize()
        k1_ms, k1_n = idx.k1_ms_total()
        idx.set_profiling(False)
        eager_ms_step = ee0.elapsed_time(ee1) / args.steps
        k1_where = "eager loop of the same K steps right after the timed loop (CUDA events around each K1 launch)"
        barrier()

    k1_rank_ms = k1_ms / max(k1_n, 1)
    k1_max_t = torch.tensor([k1_rank_ms], device=dev)
    if world > 1:                                  # the step waits for the slowest shard: its K1 is the one to subtract
        dist.all_reduce(k1_max_t, op=dist.ReduceOp.MAX)
    k1_max_ms = float(k1_max_t.item())

    # ---- end-to-end timing (host queries in, host results out, every step)
    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms_step = 1e3 * float(e2e_s.item()) / args.steps
    if rank == 0 and len(sampler.rows) < 3:      # very short runs: keep the GPU under load until a few samples exist
        t_end = time.perf_counter() + 1.0
        while len(sampler.rows) < 3 and time.perf_counter() < t_end:
            idx.search_ex(xq_dev, k, algo=args.algo)     # LOCAL search only: rank 0 is alone in this loop
            torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    barrier()

    # sanity: the timed path returned a real result (ids valid, scores descending), identical through e2e
    assert int(final_ids.min()) >= 0 and int(final_ids.max()) < n
    assert bool((final_scores[:, :-1] >= final_scores[:, 1:]).all())
    e2e_ids = graph_host.out_host["ids"] if graph_host is not None else I_pin
    assert torch.equal(e2e_ids.to(dev), final_ids), "e2e step and device step disagree"

    # ---- parity of the timed path at this N
    parity = None
    if not args.no_parity:
        pg = (sh.capture(256, k, algo=args.algo) if sh is not None else idx.capture(256, k, algo=args.algo)) \
            if use_graph else None
        if pg is None:
            class _Eager:          # same interface over the eager step
                nq = 256

                def replay(self, Q):
                    return sh.search(Q, k, algo=args.algo) if sh is not None else idx.search_ex(Q, k, algo=args.algo)
            pg = _Eager()
        parity = parity_check(torch, dist, m, idx, sh, pg, rows, n, d, k, world, rank, dev)

    if rank == 0:
        peaks = _peaks()
        flops_launch = 2.0 * nq * len(rows) * d          # algorithmic: 2*nq*N_local*d per K1 launch
        k1_avg_ms = k1_ms / max(k1_n, 1)
        achieved_tf = flops_launch / (k1_avg_ms * 1e-3) / 1e12
        traffic = None
        tf = ROOT / "profiles" / "k1_traffic.json"
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                if tj.get("rows") == len(rows) and tj.get("nq") == nq:
                    traffic = tj.get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        energy_j = None
        if clocks and clocks.get("power_w_median"):
            energy_j = clocks["power_w_median"] * k1_avg_ms * 1e-3
        line = {
            "metric": METRIC, "value": nq / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(n, d, nq, k), "sharding": f"rows sharded over {world} GPU(s)",
                       "rows_per_gpu": len(rows), "l2_policy": "inputs larger than L2 (bank shard "
                                   f"{len(rows) * d * 2 / 1e9:.2f} GB per GPU vs 126 MB L2)",
                       "search_kernel": idx.last_algo, "bank_build_s": round(build_s, 3),
                       "step": ("one CUDA graph replay of mips_search_sharded" if use_graph else "eager C-ABI calls"),
                       "exchange": (sh.exchange if sh is not None else "none"), "shards": "equal"},
            "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved_tf / peaks["tf_sustained"], "traffic": traffic,
                         "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a multi-step loop)",
                         "frac_of_burst_peak": achieved_tf / peaks["tf_burst"], "k1_ms_avg": k1_avg_ms,
                         "k1_launches_timed": k1_n, "k1_timed_in": k1_where,
                         "eager_ms_per_step": eager_ms_step, "k1_ms_avg_slowest_rank": k1_max_ms,
                         "step_minus_k1_ms": ms_step - k1_max_ms,
                         "flops_per_launch": flops_launch, "k1_energy_j_per_launch": energy_j,
                         "hbm_gbs_algorithmic": len(rows) * d * 2 / (k1_avg_ms * 1e-3) / 1e9},
            "e2e": {"value": nq / (e2e_ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nq * d * 4,
                    "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_ms_step},
            "gpu_launches": int(launches), "clocks": clocks, "parity": parity,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_sample_leg(nq, d, k, n)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if sh is not None:
        sh.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="bank rows (default: the BASELINE workload)")
    ap.add_argument("--nq", type=int, default=NQ)
    ap.add_argument("--k", type=int, default=TOPK)
    ap.add_argument("--algo", default="auto", choices=["auto", "tc", "tc128", "tc2", "tcx", "simt"], help="K1 variant (A/B runs)")
    ap.add_argument("--exchange", default=None, choices=["native", "torch", "nccl", "p2p"],
                    help="cross-GPU step: native = raw ncclAllGather issued by the C ABI on the step's stream (default); "
                         "torch = torch.distributed all-gather between C-ABI calls; p2p = peer-memory exchange")
    ap.add_argument("--e2e-graph-io", action="store_true", help="N>1 e2e: host<->device copies inside the CUDA graph")
    ap.add_argument("--no-graph", action="store_true", help="time eager C-ABI calls instead of a CUDA graph replay")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity object (profiling runs)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
