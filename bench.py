#!/usr/bin/env python
"""bench.py — the headline measurement (BASELINE.json metric: rows/s and achieved HBM GB/s per query).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2f|cfg2i|cfg4|cfg5|csv]
                    [--rows R] [--impl reference] [--no-sub]

A "step" is one pass of the hot path over one batch of synthetic input that is already resident in
HBM. The line's own workload is BASELINE.json configs[2] — the reference's hot loop,
HashAggregateExec (Main.kt:615-651): GROUP BY a 50-value Utf8 key, SUM/MIN/MAX/COUNT over 1 B rows —
STRONG-scaled: with N > 1 (torchrun, one rank per GPU) the 1 B rows are split N ways and the partial
tables are merged over NCCL inside the timed region (partition -> partial -> merge, Main.kt:1309-1325).
`sub` carries the same measurement for configs[1] (cfg2f, fused filter+project, 100 M rows per GPU, no
exchange), configs[4] (cfg5, TPC-H Q1 shape, 600 M rows split N ways) and configs[3] (cfg4, 10 M groups
over 1 B rows split N ways, NCCL all-to-all repartition), so all four are on the driver's clock; at N = 1
also `csv` (CsvDataSource.scan over 10 M records: device-resident, and end to end through the reader).

One JSON line on stdout (rank 0). `value` is device-resident throughput (CUDA events on the kernel
stream, max over ranks); `e2e` is the same metric through the C ABI with HOST buffers (pinned
host -> device copies in, result copied back, inside the timed region); `roofline` compares the
dominant kernel's algorithmic bytes/launch with the measured HBM peak; `cpu_baseline` is the CPU
oracle (a C++ restatement of the Kotlin operators — the reference itself cannot run here) timed on
this box's host cores on a bounded sample; `check` is in-bench evidence that the timed step's result
is right (sum of COUNTs = rows, group count, identical result on every rank).
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
sys.path.insert(0, os.path.join(ROOT, "query-engines_b200"))
sys.path.insert(0, ROOT)

GEN = dict(I64=1, F64=2, F64_INT=3, F64_STEP=4, UTF8=5, DATE32=6, BOOL=7)
STATES = "ALAKAZARCACOCTDEFLGAHIIDILINIAKSKYLAMEMDMAMIMNMSMOMTNENVNHNJNMNYNCNDOHOKORPARISCSDTNTXUTVTVAWAWVWIWY"
assert len(STATES) == 100


class Workload:
    def __init__(self, name, rows, dtype):
        self.name, self.rows, self.dtype = name, rows, dtype

    # overridden
    def specs(self): raise NotImplementedError
    def run(self, E, batch, dist=None): raise NotImplementedError
    def algo_bytes(self, n, out_rows): raise NotImplementedError
    kernel = ""
    describe = ""


class FilterProject(Workload):
    kernel = "kq_filter_project"

    def __init__(self, name, rows, flt):
        super().__init__(name, rows, "f64" if flt else "int64")
        self.flt = flt
        self.describe = ("SELECT a*b+c WHERE a>k AND b<m, %s columns, selectivity 0.25" % ("Float64" if flt else "Int64"))

    def specs(self):
        if self.flt:
            return [dict(kind=GEN["F64"], col_id=i, flo=0.0, fhi=1.0) for i in range(3)]
        return [dict(kind=GEN["I64"], col_id=i, ilo=0, ihi=1 << 20) for i in range(3)]

    def exprs(self, E):
        k = E.lit_f64(0.5) if self.flt else E.lit_i64(1 << 19)
        pred = E.binary("AND", E.binary("GT", E.col(0), k), E.binary("LT", E.col(1), k))
        proj = E.binary("ADD", E.binary("MUL", E.col(0), E.col(1)), E.col(2))
        return pred, [proj]

    def run(self, E, batch, dist=None):
        pred, proj = self.exprs(E)
        return E.filter_project(pred, proj, batch)

    def result_rows(self, res):
        return res.row_count()

    def algo_bytes(self, n, out_rows):
        return 24 * n + 8 * out_rows


class GroupBy(Workload):

    def __init__(self, name, rows, kind):
        super().__init__(name, rows, "f64")
        self.kind = kind
        self.describe = {"low": "GROUP BY 50-value Utf8 key: SUM/MIN/MAX/COUNT(Float64)",
                         "high": "GROUP BY Int64 key (10M distinct): SUM/MIN/MAX/COUNT(Float64)",
                         "q1": "TPC-H Q1 shape: date filter, derived projections, 2-key grouped SUMs + COUNT"}[kind]

    @property
    def kernel(self):
        # high cardinality runs the partitioned path: scatter (kq_hash_aggregate, KQ_AGG_MODE 1) + kq_agg_partition_reduce;
        # low cardinality (configs 3 and 5) the CTA-directory kernel of csrc/kq_k_agg_fe.cuh
        return "kq_hash_aggregate+kq_agg_partition_reduce" if self.kind == "high" else "kq_group_aggregate"

    def specs(self):
        if self.kind == "low":
            return [dict(kind=GEN["UTF8"], col_id=0, dict=STATES, dict_width=2), dict(kind=GEN["F64"], col_id=1, flo=0.0, fhi=1000.0)]
        if self.kind == "high":
            return [dict(kind=GEN["I64"], col_id=0, ilo=0, ihi=10_000_000), dict(kind=GEN["F64"], col_id=1, flo=0.0, fhi=1000.0)]
        return [dict(kind=GEN["DATE32"], col_id=0, ilo=8036, ihi=10562),
                dict(kind=GEN["UTF8"], col_id=1, dict="ANR", dict_width=1), dict(kind=GEN["UTF8"], col_id=2, dict="FO", dict_width=1),
                dict(kind=GEN["F64_INT"], col_id=3, ilo=1, ihi=51), dict(kind=GEN["F64"], col_id=4, flo=900.0, fhi=105000.0),
                dict(kind=GEN["F64_STEP"], col_id=5, ilo=0, ihi=11, fhi=100.0), dict(kind=GEN["F64_STEP"], col_id=6, ilo=0, ihi=9, fhi=100.0)]

    def make(self, E):
        if self.kind == "q1":
            one = E.lit_f64(1.0)
            disc_price = E.binary("MUL", E.col(4), E.binary("SUB", one, E.col(5)))
            charge = E.binary("MUL", disc_price, E.binary("ADD", one, E.col(6)))
            pred = E.binary("LE", E.col(0), E.lit_date32(10471))
            return E.HashAggregate([E.col(1), E.col(2)], [("SUM", E.col(3)), ("SUM", E.col(4)), ("SUM", disc_price),
                                                        ("SUM", charge), ("COUNT", E.lit_i64(1))], pred=pred, expected_groups=6)
        # the planner's cardinality estimate (dictionary sizes of the synthetic table): a sizing hint, not a limit
        hint = 10_000_000 if self.kind == "high" else 50
        v = E.col(1)
        return E.HashAggregate([E.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], expected_groups=hint)

    def run(self, E, batch, dist=None):
        agg = self.make(E)
        agg.update(batch)
        if dist is not None and dist["world"] > 1:
            if self.kind == "high":
                agg.repartition_alltoall()
            else:
                agg.merge_allreduce()
        return agg.finalize()

    def result_rows(self, res):
        return res.row_count()

    def algo_bytes(self, n, out_rows):
        per_row = {"low": 14, "high": 16, "q1": 46}[self.kind]
        return per_row * n + 40 * out_rows


class CsvScan(Workload):
    """CsvDataSource.scan (Main.kt:276-357) on the device — CSV text -> Utf8 columns."""

    kernel = "k_csv_fields"

    def __init__(self, name, rows):
        super().__init__(name, rows, "u8")
        self.describe = "CsvDataSource.scan: CSV text (6 columns, ~39 B/record, quoted fields with delimiters/line breaks) -> 6 Utf8 columns"
        self.text = None

    def make_text(self, rows, seed=7):
        """`rows` records: a 100k-record block of tests/csv_cases.synthetic repeated (ids repeat; synthetic data)."""
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from csv_cases import synthetic
        blk_rows = min(rows, 100_000)
        t = synthetic(blk_rows, seed=seed)
        head, body = t.split(b"\n", 1)
        # whole blocks only, so that every record is complete; the tail is cut from a fresh block at a record boundary
        reps, rest = divmod(rows, blk_rows)
        tail = b""
        if rest:
            tail = synthetic(rest, seed=seed).split(b"\n", 1)[1]
        return head + b"\n" + body * reps + tail

    def prepare(self, E, ctx, row0, row1):
        """The file's bytes resident in HBM (uploaded as an Int64 column; padded with empty lines, which rule C3 skips)."""
        import numpy as np
        import pyarrow as pa
        import kqgpu
        self.text = self.make_text(row1 - row0)
        pad = (-len(self.text)) % 8
        self.dev = kqgpu.Column.from_arrow(ctx, pa.array(np.frombuffer(self.text + b"\n" * pad, dtype=np.int64)))
        self.dev_ptr = self.dev.device_ptrs()[2]
        self.nbytes = len(self.text) + pad
        return self.dev

    def run(self, E, batch, dist=None):
        return E.csv_scan_ptr(self.dev_ptr, self.nbytes, True)

    def result_rows(self, res):
        sizes = [res.field(i).sizes() for i in range(res.num_columns())]
        self.out_bytes = sum(nb + 4 * (n + 1) for n, nb, _ in sizes)           # Arrow data + offsets buffers written
        return res.row_count()

    def algo_bytes(self, n, out_rows):
        return self.nbytes + getattr(self, "out_bytes", 0)         # the text read once + the Arrow buffers written once


WORKLOADS = {
    "cfg2f": lambda rows: FilterProject("cfg2f", rows or 100_000_000, True),
    "cfg2i": lambda rows: FilterProject("cfg2i", rows or 100_000_000, False),
    "cfg3": lambda rows: GroupBy("cfg3", rows or 1_000_000_000, "low"),
    "cfg4": lambda rows: GroupBy("cfg4", rows or 1_000_000_000, "high"),
    "cfg5": lambda rows: GroupBy("cfg5", rows or 600_037_902, "q1"),
    "csv": lambda rows: CsvScan("csv", rows or 10_000_000),
}


def shard_range(rank, rows_per_gpu):
    """Rows [begin, end) of the global table that `rank` owns: contiguous, equal-sized shards (weak scaling:
    the global table has world * rows_per_gpu rows). The generator is keyed by the global row index, so the
    union of all shards is the same table at every GPU count."""
    return rank * rows_per_gpu, (rank + 1) * rows_per_gpu


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for t, l in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for _, l in self.lines]
        sm, mx, reasons = [], None, set()
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload, kernel):
    """dram bytes/launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload, {}).get(kernel)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------
def cpu_reference(wl, sample_rows, threads, seed=42):
    """Time the CPU oracle (port of the Kotlin operators) on `sample_rows` rows of the same workload."""
    from oracle import oracle as O
    O.build()
    if isinstance(wl, CsvScan):                # single-threaded, like the reference's ReaderIterator (Main.kt:204-273)
        text = wl.make_text(sample_rows)
        t0 = time.perf_counter()
        out_rows = O.csv_scan(text, True).row_count()
        dt = time.perf_counter() - t0
        return sample_rows / dt, dt, out_rows
    batch = O.generate(wl.specs(), seed, 0, sample_rows)

    t0 = time.perf_counter()
    if isinstance(wl, FilterProject):
        pred, proj = wl.exprs(O)
        out_rows = O.filter_project_mt(pred, proj, batch, threads)
    else:
        if wl.kind == "q1":
            one = O.lit_f64(1.0)
            dp = O.binary("MUL", O.col(4), O.binary("SUB", one, O.col(5)))
            ch = O.binary("MUL", dp, O.binary("ADD", one, O.col(6)))
            pred = O.binary("LE", O.col(0), O.lit_date32(10471))
            res = O.hashagg_mt([O.col(1), O.col(2)], [("SUM", O.col(3)), ("SUM", O.col(4)), ("SUM", dp), ("SUM", ch),
                                                      ("COUNT", O.lit_i64(1))], batch, threads, pred=pred)
        else:
            v = O.col(1)
            res = O.hashagg_mt([O.col(0)], [("SUM", v), ("MIN", v), ("MAX", v), ("COUNT", v)], batch, threads)
        out_rows = res.row_count()
    dt = time.perf_counter() - t0
    return sample_rows / dt, dt, out_rows


def cpu_sample_rows(wl):
    # sized for roughly 10-20 s of single-thread-equivalent CPU work (the oracle is row-at-a-time and boxed)
    return {"cfg2f": 24_000_000, "cfg2i": 24_000_000, "cfg3": 16_000_000, "cfg4": 8_000_000, "cfg5": 8_000_000, "csv": 4_000_000}[wl.name]


def scaling_of(wl):
    """Strong scaling (the BASELINE table split across the ranks) for the aggregate workloads, whose merge is the
    exchange step of the path; filter+project and the CSV scan have no exchange and keep their per-GPU size (weak)."""
    return "strong" if isinstance(wl, GroupBy) else "weak"


def rows_per_rank(wl, rank, world):
    """Row range [begin, end) of the global table that `rank` owns."""
    if scaling_of(wl) == "weak":
        return shard_range(rank, wl.rows)
    per = (wl.rows + world - 1) // world
    return min(wl.rows, rank * per), min(wl.rows, (rank + 1) * per)


def config_of(wl, world):
    rows_total = wl.rows if scaling_of(wl) == "strong" else wl.rows * world
    par = f"row-sharded x{world}"
    if world > 1 and isinstance(wl, GroupBy):
        par += ", NCCL all-to-all repartition" if wl.kind == "high" else ", NCCL all-gather merge of the partial tables"
    return {"workload": f"{wl.name}: {wl.describe}", "rows_total": rows_total, "rows_per_gpu": (rows_total + world - 1) // world,
            "seed": 42, "scaling": scaling_of(wl), "l2": "inputs larger than L2 (no flush needed)", "parallelism": par}


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    threads = 1 if isinstance(wl, CsvScan) else (os.cpu_count() or 1)
    sample = cpu_sample_rows(wl) // 4
    for _ in range(args.warmup):
        cpu_reference(wl, max(sample // 8, 100_000), threads)
    rows, dt = 0, 0.0
    for _ in range(args.steps):
        _, d, _ = cpu_reference(wl, sample, threads)      # d: the operator alone (the sample table is generated before the clock starts)
        rows += sample
        dt += d
    v = rows / dt
    cfg = config_of(wl, args.gpus)
    line = {"impl": "reference", "metric": "rows/s", "value": v, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": scaling_of(wl),
            "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "rows/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} rows per step of the same seeded table, already resident in host memory when the clock starts; "
                                       f"partition->partial->merge on {threads} threads"},
            "e2e": {"value": v, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU oracle = C++ restatement of the Kotlin operators (the Kotlin reference cannot be built here: no JVM); "
                    "a rate measured on a bounded sample of the workload named in config"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def result_check(wl, E, batch, res, n_local, dist, world):
    """In-bench correctness evidence for the timed step's result (the only multi-GPU check the driver's boxes see)."""
    import hashlib
    import numpy as np
    chk = {}
    if isinstance(wl, GroupBy):
        arrs = res.to_arrow()
        nk = 2 if wl.kind == "q1" else 1
        cnt = np.asarray(arrs[-1].to_numpy(zero_copy_only=False), dtype=np.int64)          # COUNT is the last aggregate of every workload
        chk["groups_rank0"] = int(len(cnt))
        local_rows_counted = int(cnt.sum())
        if wl.kind == "q1":
            pred = E.binary("LE", E.col(0), E.lit_date32(10471))
            expect_local = E.filter(pred, batch).row_count()                               # rows passing the filter, counted by another kernel
        else:
            expect_local = n_local
        # canonical bytes of the result: rows sorted by key, every column's values
        keys = [a.to_pylist() for a in arrs[:nk]]
        order = sorted(range(len(cnt)), key=lambda i: tuple(k[i] for k in keys))
        h = hashlib.sha256()
        for a in arrs:
            col = a.to_pylist()
            h.update(repr([col[i] for i in order]).encode())
        digest = h.hexdigest()[:16]
        chk["result_sha256_16"] = digest
        if dist is None:
            chk["count_sum"] = local_rows_counted
            chk["count_expected"] = int(expect_local)
            chk["count_ok"] = local_rows_counted == expect_local
        else:
            torch, td = dist["torch"], dist["td"]
            t = torch.tensor([expect_local, local_rows_counted, len(cnt)], device="cuda", dtype=torch.int64)
            g = [torch.zeros_like(t) for _ in range(world)]
            td.all_gather(g, t)
            exp_total = int(sum(int(x[0]) for x in g))
            if wl.kind == "high":           # repartitioned: every key lives on exactly one rank, concatenation = answer
                chk["count_sum"] = int(sum(int(x[1]) for x in g))
                chk["groups_total"] = int(sum(int(x[2]) for x in g))
            else:                           # all-gather merge: every rank holds the full result
                chk["count_sum"] = local_rows_counted
                d8 = torch.tensor(list(bytes.fromhex(digest)), device="cuda", dtype=torch.uint8)
                dg = [torch.zeros_like(d8) for _ in range(world)]
                td.all_gather(dg, d8)
                chk["all_ranks_identical"] = all(bool((x == d8).all()) for x in dg)
            chk["count_expected"] = exp_total
            chk["count_ok"] = chk["count_sum"] == exp_total
        if wl.kind == "low":
            chk["groups_ok"] = chk["groups_rank0"] == 50
    elif isinstance(wl, FilterProject):
        chk["rows_out_rank0"] = res.row_count()
    return chk


def measure(kqgpu, ctx, E, wl, args, dist, rank, world, steps, warmup, want_e2e, want_cpu):
    """One workload: generate this rank's shard, warm up, time `steps` steps on the device, check the result."""
    row0, row1 = rows_per_rank(wl, rank, world)
    n = row1 - row0
    batch = wl.prepare(E, ctx, row0, row1) if hasattr(wl, "prepare") else E.generate(wl.specs(), 42, row0, row1)
    ctx.sync()

    def barrier():
        ctx.sync()
        if dist:
            dist["td"].barrier()
            dist["torch"].cuda.synchronize()

    out_rows = 0
    keep = None
    for _ in range(warmup):
        keep = wl.run(E, batch, dist)      # same ownership pattern as the timed loop: the previous result lives until the next exists
        out_rows = wl.result_rows(keep)
    barrier()

    sampler = ClockSampler(ctx.device) if rank == 0 else None
    launches0 = ctx.launch_count()
    t_wall0 = time.time()
    ctx.timer_begin()
    for _ in range(steps):
        keep = wl.run(E, batch, dist)          # inputs (2.4-28 GB) are far larger than L2: no flush needed
    ms = ctx.timer_end()
    barrier()
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None
    out_rows = wl.result_rows(keep)
    check = result_check(wl, E, batch, keep, n, dist, world)
    del keep
    if dist:
        t = dist["torch"].tensor([ms], device="cuda")
        dist["td"].all_reduce(t, op=dist["td"].ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / steps
    cfg = config_of(wl, world)
    rows_total = cfg["rows_total"]
    value = rows_total / (ms_per_step * 1e-3)

    e2e = None
    if want_e2e:
        e2e = run_e2e(kqgpu, ctx, E, wl, batch, dist, min(steps, 5), world, n, rows_total)

    peak, peak_src = measured_peak()
    # roofline of the dominant kernel on THIS rank's shard (per launch), against the per-GPU peak
    algo = wl.algo_bytes(n, out_rows)
    achieved = algo / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(wl.name, wl.kernel), "kernel": wl.kernel, "algorithmic_bytes_per_launch": algo,
                "peak_source": peak_src, "frac_of_8TBps": achieved / 8000.0,
                "how": "algorithmic bytes of one rank's step / (CUDA-event time of the timed region / steps, max over ranks) on the kernel "
                       "stream; at N>1 the step also holds the NCCL merge, so this is a lower bound for the kernel"}
    cpu = None
    if want_cpu and rank == 0 and world == 1:
        threads = 1 if isinstance(wl, CsvScan) else (os.cpu_count() or 1)      # the reference's CSV reader is one thread (Main.kt:204-273)
        sample = cpu_sample_rows(wl)
        v1, dt1, _ = cpu_reference(wl, sample // 8, 1)
        vn, dtn, _ = cpu_reference(wl, sample, threads)
        cpu = {"value": vn, "unit": "rows/s", "cores": threads, "kind": "port",
               "sample": f"{sample} rows of the same seeded table, partition->partial->merge on {threads} threads ({dtn:.1f} s); "
                         f"1 thread on {sample // 8} rows: {v1:.3e} rows/s",
               "value_1thread": v1,
               "note": "C++ restatement of the Kotlin operators (boxed, row-at-a-time); the Kotlin reference cannot run here (no JVM)"}
    del batch
    return {"value": value, "output_rows_rank0": out_rows, "ms_per_step": ms_per_step, "scaling": scaling_of(wl), "dtype": wl.dtype, "config": cfg, "roofline": roofline,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "check": check, "steps": steps, "warmup": warmup}


def bind_to_gpu_numa(local_rank):
    """Run this rank (and first-touch its pinned staging buffers) on the CPUs nearest to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs nearest to GPU {local_rank}"
    except Exception as e:       # no NVML / restricted container: leave the affinity alone
        return f"unchanged ({type(e).__name__})"
    return "unchanged"


SUBS = {"cfg3": ["cfg2f", "cfg5", "cfg4"]}


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="rows of the table (strong-scaled workloads: in total; weak: per GPU); default: the BASELINE config size")
    ap.add_argument("--impl", default="kqgpu", choices=["kqgpu", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="skip the secondary workloads reported under `sub`")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "kqgpu" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload](args.rows)

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import kqgpu
    if kqgpu.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU oracle)")
    numa = bind_to_gpu_numa(local_rank)

    dist = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local_rank)
        td.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = {"td": td, "torch": torch, "world": world, "rank": rank}

    ctx = kqgpu.Context(local_rank)
    E = kqgpu.Engine(ctx)
    if world > 1:
        import ctypes
        idbuf = ctypes.create_string_buffer(kqgpu.COMM_ID_BYTES)
        if rank == 0:
            ctx.check(kqgpu.lib().kq_comm_unique_id(ctx.h, idbuf))
        t = dist["torch"].frombuffer(bytearray(idbuf.raw), dtype=dist["torch"].uint8).cuda()
        dist["td"].broadcast(t, 0)
        idbytes = bytes(t.cpu().numpy().tobytes())
        ctx.check(kqgpu.lib().kq_comm_init(ctx.h, idbytes, rank, world))

    m = measure(kqgpu, ctx, E, wl, args, dist, rank, world, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline)
    sub = {}
    if not args.no_sub and not args.rows:
        for name in SUBS.get(args.workload, []):
            swl = WORKLOADS[name](0)
            r = measure(kqgpu, ctx, E, swl, args, dist, rank, world, max(3, min(args.steps, 10)), 3, not args.no_e2e, False)
            sub[name] = {k: r[k] for k in ("value", "ms_per_step", "scaling", "dtype", "config", "roofline", "e2e", "gpu_launches", "check", "steps", "output_rows_rank0")}
            sub[name]["unit"] = "rows/s"
        if args.workload == "cfg3" and world == 1:
            # the scan side on the same clock: CsvDataSource.scan over 10 M records, device-resident and end to
            # end through the reader. Last and guarded: whatever happens here, the line above is printed.
            try:
                r = measure(kqgpu, ctx, E, WORKLOADS["csv"](0), args, dist, rank, world, 5, 3, not args.no_e2e, False)
                sub["csv"] = {k: r[k] for k in ("value", "ms_per_step", "scaling", "dtype", "config", "roofline", "e2e", "gpu_launches", "check", "steps", "output_rows_rank0")}
                sub["csv"]["unit"] = "rows/s"
            except Exception as e:      # noqa: BLE001
                sub["csv"] = {"error": repr(e)[:300]}

    if rank == 0:
        line = {"metric": "rows/s", "value": m["value"], "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": m["scaling"], "vs_baseline": None, "dtype": m["dtype"],
                "data": "synthetic", "config": m["config"], "roofline": m["roofline"], "cpu_baseline": m["cpu_baseline"], "e2e": m["e2e"],
                "gpu_launches": m["gpu_launches"], "clocks": m["clocks"], "check": m["check"], "output_rows_rank0": m["output_rows_rank0"],
                "cpu_affinity": numa, "sub": sub}
        print(json.dumps(line), flush=True)
    if dist:
        dist["td"].destroy_process_group()


def run_e2e(kqgpu, ctx, E, wl, batch, dist, steps, world, n_local, rows_total):
    """Same step, but inputs start in pinned HOST memory and the result ends in host memory.

    filter+project: one kq_filter_project_host call (chunked H2D / kernel / D2H overlap inside the library).
    group-by: the drain loop of HashAggregateExec over host batches (kq_column_upload -> kq_hashagg_update per
    batch of 32 Mi rows, Main.kt:617-634), merge across ranks, finalize, result batch copied back."""
    import ctypes as C
    L = kqgpu.lib()
    if isinstance(wl, CsvScan):
        # the file's bytes in pinned host memory -> kq_csv_reader (H2D inside, overlapped) -> row counts and buffer sizes read back
        nb = len(wl.text)
        hp = ctx.host_alloc(nb)
        C.memmove(hp, wl.text, nb)

        def one_csv():
            d2h = rows = 0
            for res in E.csv_batches(hp, True, nbytes=nb):                          # Sequence<RecordBatch>: one batch per 64 MiB piece
                sizes = [res.field(i).sizes() for i in range(res.num_columns())]    # the step's result read back: rows and bytes per column
                d2h += 8 + 24 * len(sizes)
                rows += res.row_count()
            assert rows == n_local, (rows, n_local)
            return d2h, rows
        d2h, _ = one_csv()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(steps):
            d2h, _ = one_csv()
        ctx.sync()
        dt = time.perf_counter() - t0
        ctx.host_free(hp)
        return {"value": rows_total * steps / dt, "unit": "rows/s", "h2d_bytes_per_step": nb, "d2h_bytes_per_step": d2h, "steps": steps,
                "ms_per_step": dt / steps * 1e3,
                "how": "CSV text in pinned host memory -> kq_csv_reader (64 MiB pieces, the H2D copy of a piece under the scan of the one before) -> row count and column sizes of every batch read back, wall clock around synchronised steps"}
    cols = [batch.field(i) for i in range(batch.num_columns())]
    host = []
    h2d = 0
    for c in cols:
        n, nb, nn = c.sizes()
        t = c.type()
        ptr_d = ctx.host_alloc(max(nb, 1))
        ptr_o = ctx.host_alloc((n + 1) * 4) if t == kqgpu.UTF8 else None
        ptr_v = ctx.host_alloc((n + 7) // 8 + 8) if nn > 0 else None
        ctx.check(L.kq_column_download(ctx.h, c.h, C.c_void_p(ptr_v) if ptr_v else None,
                                       C.c_void_p(ptr_o) if ptr_o else None, C.c_void_p(ptr_d)))
        host.append((t, n, ptr_v, ptr_o, ptr_d, nb))
        h2d += nb + ((n + 1) * 4 if ptr_o else 0) + ((n + 7) // 8 if ptr_v else 0)
    n = n_local
    outs = []
    if isinstance(wl, FilterProject):
        pred, proj = wl.exprs(E)
        outs = [(ctx.host_alloc(max(n, 1) * 8), None) for _ in proj]

        def one():
            m = E.filter_project_host(pred, proj, [(t, pv, pd) for (t, _, pv, po, pd, nb) in host], n, outs)
            return m * 8 * len(proj), m
    else:
        CH = 32 << 20

        def one():
            agg = wl.make(E)
            for r0 in range(0, n, CH):
                r1 = min(n, r0 + CH)
                up = []
                for (t, _, pv, po, pd, nb) in host:
                    out = C.c_void_p()
                    w = {kqgpu.F64: 8, kqgpu.I64: 8, kqgpu.DATE32: 4}.get(t, 0)
                    if t == kqgpu.UTF8:     # a slice of an Arrow Utf8 vector: offsets window, same data buffer
                        ctx.check(L.kq_column_upload(ctx.h, t, r1 - r0, C.c_void_p(pv + r0 // 8) if pv else None, C.c_void_p(po + 4 * r0),
                                                     C.c_void_p(pd), nb, C.byref(out)))
                    else:
                        ctx.check(L.kq_column_upload(ctx.h, t, r1 - r0, C.c_void_p(pv + r0 // 8) if pv else None, None,
                                                     C.c_void_p(pd + w * r0), 0, C.byref(out)))
                    up.append(kqgpu.Column(ctx, out))
                agg.update(kqgpu.RecordBatch.from_columns(ctx, up))
            if dist is not None and dist["world"] > 1:
                agg.repartition_alltoall() if wl.kind == "high" else agg.merge_allreduce()
            arrs = agg.finalize().to_arrow()             # device -> host copy of the step's result
            return sum(a.nbytes for a in arrs), len(arrs[0]) if arrs else 0

    d2h, _ = one()                        # warm-up
    if dist:
        dist["td"].barrier()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        d2h, _ = one()
    ctx.sync()
    dt = time.perf_counter() - t0
    if dist:
        t = dist["torch"].tensor([dt], device="cuda")
        dist["td"].all_reduce(t, op=dist["td"].ReduceOp.MAX)
        dt = float(t.item())
    for (t, n_, pv, po, pd, nb) in host:
        for p in (pv, po, pd):
            if p:
                ctx.host_free(p)
    for (pd, pv) in outs:
        ctx.host_free(pd)
    return {"value": rows_total * steps / dt, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "steps": steps, "ms_per_step": dt / steps * 1e3,
            "how": ("kq_filter_project_host: pinned host buffers -> chunked H2D / fused kernel / D2H overlap -> host result buffers"
                    if isinstance(wl, FilterProject) else
                    "pinned host buffers -> kq_column_upload + kq_hashagg_update per 32 Mi-row batch -> finalize -> result downloaded") +
                   ", wall clock around synchronised steps"}


if __name__ == "__main__":
    main()
