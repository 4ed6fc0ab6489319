#!/usr/bin/env python
"""bench.py — alignments/s and CIGAR ops/s of the signal-extraction hot path on N B200s.

A step = one pass of the hot path (kernels 0-4 + the result header) over one batch: BASELINE.json configs[1], 1M simulated
HiFi molecules (~1.04M alignment records, ~5 % carrying SA tags), default parameters + -p 0.8.

  value        alignments/s, inputs resident in HBM, every step timed on the device with the library's CUDA events
               on the stream its kernels run on (L2 flushed before each step); max over ranks, summed over ranks.
  e2e          the same metric through the public API with HOST buffers: pinned SoA views -> exlr_submit (H2D + kernels +
               results copied back behind them) -> exlr_wait, sub-batches pipelined over several streams.
  roofline     the byte-dominant kernel (1a, the streaming CIGAR screen) against the measured HBM copy peak, plus a table
               of EVERY kernel of the step (time, algorithmic bytes, fraction of the peak) naming the time-dominant one.
  cpu_baseline the CPU oracle (a port of the reference loop; the Rust reference cannot be built here) on the same batch.
  strong_c5    (N > 1) BASELINE.json configs[4]: ONE 6M-record set cut into 64k-record batches dealt round robin to the
               ranks, the same two measurements, next to rank 0 running the whole set alone.

`--impl reference` times that CPU port with every host core (records sharded over threads) instead of the GPU.
Multi-GPU: reads are independent, so each rank processes its own shard (weak scaling, no collective on the data
path); torch.distributed is used only for the barrier and the max-over-ranks reduction of the time.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """dram bytes per launch of every kernel from the committed ncu --set full capture (profiles/traffic.json), if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_ev = index, [], set(), None, threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}
        while not self._stop_ev.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._stop_ev.set()
        if self.is_alive():
            self.join(timeout=1)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def make_workload(config_id, scale, rank, world=1, strong=False):
    """weak scaling (default): every rank gets its own batch of the config's size (seed + rank).
    strong: ONE dataset, cut into 64k-record batches that go round robin to the ranks; rank r keeps its share."""
    from excord_lr_b200 import shard, synth
    from excord_lr_b200.batch import HostBatch
    c = synth.CONFIGS[config_id]
    n = max(1, int(c["n"] * scale))
    if strong:
        hb = synth.generate(c["profile"], c["seed"], n, c["chr20"], c["ultra"] if scale >= 1 else 0)
        if world == 1:
            return hb, c, hb
        mine = shard.rank_batches(shard.plan_batches(hb.n_reads, 65536), rank, world)
        return HostBatch.concat([hb.slice(a, b) for _, a, b in mine]), c, hb
    hb = synth.generate(c["profile"], c["seed"] + 7919 * rank, n, c["chr20"], c["ultra"] if scale >= 1 else 0)
    return hb, c, hb


def config_dict(c, hb):
    """The workload as both arms state it (identical keys and values in the GPU arm and in `--impl reference`)."""
    return {"workload": c["name"], "records_per_gpu": hb.n_reads, "cigar_ops_per_gpu": hb.n_ops, "sa_bytes_per_gpu": hb.n_sa_bytes,
            "params": c["params"], "parallelism": "record shards, one per GPU, no collective",
            "l2": "GPU arm: the timed steps go round robin over several device-resident copies of the batch (together several times the 126 MB L2, "
                  "so no step finds its inputs cached; no flush kernel inside the timed region); the single_batch figure flushes L2 (256 MB read) "
                  "before every step; CPU arm: not applicable"}


def split_for_pipeline(hb, parts):
    n = hb.n_reads
    bounds = [n * i // parts for i in range(parts + 1)]
    return [hb.slice(bounds[i], bounds[i + 1]) for i in range(parts)]


def run_reference(args, rank, world):
    """CPU arm: the oracle port of the reference loop, sharded over every host core."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_c
    from excord_lr_b200.batch import ExlrParams
    if rank != 0:
        return
    hb, c, _ = make_workload(args.config, args.scale, 0)
    p = ExlrParams.make(**c["params"])
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        oracle_c.run_sharded(hb, p, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rs = oracle_c.run_sharded(hb, p, cores)
    dt = (time.perf_counter() - t0) / args.steps
    assert all(r.status == 0 for r in rs)
    v = hb.n_reads / dt
    line = {"impl": "reference", "metric": "alignments_per_sec", "value": v, "unit": "alignments/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "cigar_ops_per_sec": hb.n_ops / dt,
            "config": config_dict(c, hb),
            "cpu_baseline": {"value": v, "unit": "alignments/s", "cores": cores, "kind": "port",
                             "sample": f"whole workload ({hb.n_reads} records) per step, records sharded over {cores} threads; "
                                       "the reference's own loop is single-threaded (src/main.rs:158)"},
            "e2e": {"value": v, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


class Dist:
    """barrier / reductions over the ranks (torch.distributed over NCCL; nothing on the data path)."""

    def __init__(self, world, local_rank):
        import torch
        self.torch, self.world = torch, world
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":      # NCCL would print its version banner on stdout, next to the JSON line
                os.environ["NCCL_DEBUG"] = "WARN"
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op={"max": self.dist.ReduceOp.MAX, "sum": self.dist.ReduceOp.SUM, "min": self.dist.ReduceOp.MIN}[op])
        return t.tolist()

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


class Flusher:
    def __init__(self, torch):
        self.torch = torch
        self.buf = torch.zeros(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def __call__(self):
        self.buf.sum()                  # READ 256 MB: evicts the batch from L2 and leaves only clean lines behind
        self.torch.cuda.synchronize()


def measure_resident(big, flush, steps, warmup, D):
    """K timed steps on the device-resident batch: (sum of step ms, launches, lines per step)."""
    def step():
        flush()
        big.submit_resident()
        r = big.wait_resident()
        return r, big.timing()
    for _ in range(warmup):
        r, _t = step()
    assert r.status == 0, f"status {r.status}"
    D.barrier()
    ms, launches = [], 0
    for _ in range(steps):
        r, t = step()
        ms.append(t.kernels_ms + t.d2h_ms)                                # first kernel start -> result header on the host
        launches += t.launches
    D.barrier()
    return float(np.sum(ms)), launches, r.n_events, step


def measure_pipelined(batches, steps, warmup, D):
    """K timed steps over several device-resident copies of the batch, round robin, each copy on its own streams: a copy is waited
    (result header on the host) right before it is submitted again, so up to len(batches) steps are in flight and the tail of one
    (the latency-bound ordering kernels) runs under the head of the next -- how a stream of batches goes through one GPU.
    Timed on the device: events recorded while the GPU is idle before the first submit and after the last result is on the host.
    -> (ms for the K steps, launches, lines per step)"""
    torch = D.torch
    nb = len(batches)

    def run(k):
        busy, r = [False] * nb, None
        for s in range(k):
            b = batches[s % nb]
            if busy[s % nb]:
                r = b.wait_resident(); assert r.status == 0, f"status {r.status}"
            b.submit_resident(); busy[s % nb] = True
        for i, b in enumerate(batches):
            if busy[i]:
                r = b.wait_resident(); assert r.status == 0, f"status {r.status}"
        return r.n_events
    run(max(warmup, 3 * nb))                                               # (every copy has replayed its graph at least once)
    D.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_events = run(steps)
    e1.record()
    torch.cuda.synchronize()
    D.barrier()
    return float(e0.elapsed_time(e1)), int(batches[0].timing().launches) * steps, n_events


def measure_e2e(ex, hb, n_parts, steps, warmup, n_events, D):
    """The same steps through exlr_submit / exlr_wait the way a streaming caller (the CLI) uses them: a ring of sub-batches, each
    waited (events + line offsets on the host) right before its buffers are re-submitted with the next step's records, so the
    H2D engine never idles between steps.  Every step's inputs cross PCIe and every step's result is read back.
    -> (seconds, h2d bytes per step, d2h bytes per step) -- the byte counts are the library's own (exlr_timing)."""
    torch = D.torch
    parts = split_for_pipeline(hb, max(1, n_parts))
    pb = [ex.batch_for(h) for h in parts]                                  # pinned views already hold the packed records

    def run(k):
        tot, busy = 0, [False] * len(pb)
        for _ in range(k):
            for i, b in enumerate(pb):
                if busy[i]:
                    tot += b.wait(copy=False).n_events
                b.submit()
                busy[i] = True
        for i, b in enumerate(pb):
            if busy[i]:
                tot += b.wait(copy=False).n_events
        return tot
    ne = run(warmup)
    assert ne == n_events * warmup, f"pipelined run produced {ne} lines, resident run {n_events} per step"
    D.barrier()
    t0 = time.perf_counter()
    ne = run(steps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    D.barrier()
    assert ne == n_events * steps
    h2d = sum(int(b.timing().h2d_bytes) for b in pb)
    d2h = sum(int(b.timing().d2h_bytes) for b in pb)
    for b in pb:
        b.free()
    return dt, h2d, d2h, len(parts)


def pcie_probe(D, concurrent):
    """Pinned host -> device copy rate of this rank: one 256 MB copy, best of 10.  concurrent=True: every rank copies at the same
    time (barrier-aligned rounds), which is what the ranks' e2e loops do to the box's host memory and PCIe root."""
    torch = D.torch
    pin = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
    dev = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rates = []
    for _ in range(10):
        if concurrent:
            D.barrier()
        e0.record(); dev.copy_(pin, non_blocking=True); e1.record(); torch.cuda.synchronize()
        rates.append((256 << 20) / (e0.elapsed_time(e1) / 1e3) / 1e9)
    del pin, dev
    rates.sort()
    return rates[-1] if not concurrent else rates[len(rates) // 2]       # alone: best; together: the median round


def kernel_table(hb, p, solo, cnt, n_events, peak, traffic):
    """Every kernel of a step: same-stream stage time (the kernel alone on the GPU, CUDA events between the kernels), the
    algorithmic bytes it has to move, and what fraction of the HBM peak that is."""
    R, C, A = hb.n_reads, hb.n_ops, hb.n_sa_bytes
    has_sa = hb.sa_kind != 0
    S = int(has_sa.sum())
    sa_ops = int(np.diff(hb.cigar_off.astype(np.int64))[has_sa].sum())
    claimed = cnt["claimed_short"] + cnt["claimed_warp"] + cnt["claimed_long"]
    raw, sa_ev, E = cnt["raw_events"], cnt["sa_events"], n_events
    rows = [("k0_classify", solo["classify_ms"], 8 * R + 4 * R + 4 * S, "flag 2 + mapq 1 + sa_kind 1 + tid 4 read, 4 B/record + 4 B/SA record written"),
            ("k1a_screen", solo["screen_ms"], 4 * C, "every CIGAR op read once")]
    if not p.split_only:
        rows.append(("k1b_claim+k1b_walk", solo["cigar_ms"] - solo["screen_ms"], 2048 * cnt["flagged_steps"] + 19 * claimed + 32 * raw,
                     "the flagged 512-op steps read again, 19 B per claimed record, 32 B per raw event written"))
    rows += [("k3a_sa_cigar", solo["sa_cigar_ms"], 4 * sa_ops + 12 * S + 32 * S, "4 B/op of the SA records' own CIGARs, 32 B summary written"),
             ("k3b_sa_events", solo["sa_parse_ms"], A + 32 * S + 16 * S + 48 * sa_ev, "SA bytes + 32 B summary + 16 B/record read, 48 B/line written"),
             ("k4a_line_scan", solo["scan_ms"], 4 * R + R // 8 + 8 * claimed + 4 * R, "4 B/record + claim bit + 8 B/claimed record read, 4 B/record written"),
             ("k4b_place", solo["place_ms"], 32 * raw + 48 * sa_ev + 48 * E + 128, "raw + SA events read, 48 B/line written; its last CTA stores the 128-byte result header")]
    out = []
    for name, ms, b, what in rows:
        if ms <= 0:
            continue
        tr = (traffic.get(name.split("+")[0].split("(")[0]) or {}).get("dram_bytes_per_launch")
        out.append({"name": name, "avg_ms": ms, "algorithmic_bytes": int(b), "achieved_gbs": b / (ms / 1e3) / 1e9,
                    "frac": b / (ms / 1e3) / 1e9 / peak, "traffic": tr, "bytes": what})
    return out


def strong_c5(args, ex_opts, D, rank, local_rank, world, peak):
    """BASELINE.json configs[4] (6M HiFi records) sharded round robin in 64k-record batches over the ranks: fixed total work."""
    from excord_lr_b200 import api
    from excord_lr_b200.batch import ExlrParams
    torch = D.torch
    hb, c, whole = make_workload(4, args.strong_scale, rank, world, strong=True)
    p = ExlrParams.make(**c["params"])
    steps, warmup = min(args.steps, 5), 3
    flush = Flusher(torch)

    def one(h, barrier_obj):
        ex = api.Extractor(p, h.ref_names, local_rank)
        for k, v in ex_opts:
            ex.set_option(k, v)
        ex.set_option(api.EXLR_OPT_STAGE_TIMING, 0)
        big = ex.batch_for(h)
        big.upload()
        tot, _l, ne, _ = measure_resident(big, flush, steps, warmup, barrier_obj)
        big.free()
        dt, h2d, d2h, _n = measure_e2e(ex, h, args.pipeline_parts, steps, warmup, ne, barrier_obj)
        ex.close()
        return tot, dt, h2d, d2h
    tot, dt, h2d, d2h = one(hb, D)
    tot_max, e2e_max = D.reduce([tot, dt * 1e3], "max")
    R_all = whole.n_reads
    out = {"workload": c["name"], "records_total": R_all, "cigar_ops_total": whole.n_ops, "batch_records": 65536, "scaling": "strong",
           "steps": steps, "warmup": warmup,
           "value": R_all / (tot_max / 1e3) * steps, "ms_per_step": tot_max / steps,
           "e2e": {"value": R_all / (e2e_max / 1e3) * steps, "ms_per_step": e2e_max / steps,
                   "h2d_bytes_per_step_rank0": h2d, "d2h_bytes_per_step_rank0": d2h}}
    if world > 1:
        # the same set on ONE GPU (rank 0, the others wait): what the N-GPU numbers are to be compared with
        class Solo:
            torch = D.torch

            def barrier(self):
                torch.cuda.synchronize()
        if rank == 0:
            t1, d1, _h, _d = one(whole, Solo())
            out["single_gpu"] = {"value": R_all / (t1 / 1e3) * steps, "ms_per_step": t1 / steps,
                                 "e2e_value": R_all / d1 * steps, "e2e_ms_per_step": d1 * 1e3 / steps}
            out["efficiency"] = out["value"] / (world * out["single_gpu"]["value"])
            out["e2e_efficiency"] = out["e2e"]["value"] / (world * out["single_gpu"]["e2e_value"])
        D.barrier()
    return out


def e2e_bam(args, pcie_gbs):
    """BAM bytes in -> event bytes out, through the CLI (excord_lr_b200/host/excord-lr-b200): the file is read into pinned memory,
    inflated and walked on the GPU (exlr_bam_*), and the lines come back formatted.  Workload: BASELINE.json configs[1] molecules
    with 15 kb of random SEQ/QUAL per record (what makes a real HiFi BAM big); the same run with --host-reader (zlib on every host
    core) beside it.  stream_s is the CLI's own `stream` figure (after process setup: CUDA context, buffers), wall_s the whole process."""
    import re
    import subprocess
    import tempfile
    from excord_lr_b200 import bamio, synth
    exe = os.path.join(ROOT, "excord_lr_b200", "host", "excord-lr-b200")
    if not os.path.exists(exe):
        return {"unavailable": "host/excord-lr-b200 not built"}
    hb = synth.config(1, args.bam_scale)
    cores = os.cpu_count() or 1
    out = {"workload": f"c2_hifi molecules x {args.bam_scale}, 15 kb random SEQ/QUAL per record, BGZF level 1", "records": hb.n_reads}
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
        bam, txt = os.path.join(d, "x.bam"), os.path.join(d, "x.txt")
        bamio.write_bam(hb, bam, ref_lens=synth.ref_lens(), seq_len=15000, random_seq=True)
        size = os.path.getsize(bam)
        out["bam_bytes"] = size
        ref = None
        for key, extra in (("gpu_decoder", []), ("host_reader", ["--host-reader"])):
            # stream_s: the CLI's own `stream` figure with every batch in place before the input is read (--eager-alloc) -- the rate a
            # large file sees; wall_s: the whole process as a user starts it (CUDA context, batches allocated while the first ones
            # already work), best of 3 each
            best, wall, detail = None, None, ""
            for eager in (True, True, True, False, False, False):
                t0 = time.perf_counter()
                r = subprocess.run([exe, "-b", bam, "-o", txt, "-p", "0.8", "-t", str(cores), "--stats"] + extra + (["--eager-alloc"] if eager else []),
                                   capture_output=True, text=True)
                w = time.perf_counter() - t0
                if r.returncode != 0:
                    return {"unavailable": r.stderr[-300:]}
                if not eager:
                    wall = w if wall is None else min(wall, w)
                    continue
                m = re.search(r"stream ([0-9.]+) s", r.stderr)
                st = float(m.group(1)) if m else w
                if best is None or st < best:
                    best = st
                    dm = re.search(r"GPU BAM decoder: (.*)", r.stderr)
                    detail = dm.group(1) if dm else ""
            got = open(txt, "rb").read()
            if ref is not None and got != ref:
                return {"unavailable": "the GPU decoder and the host reader wrote different files"}
            ref = got
            out[key] = {"stream_s": best, "wall_s_with_process_setup": wall, "alignments_per_sec": hb.n_reads / best, "bam_gbs": size / best / 1e9,
                        "frac_of_pcie_h2d": size / best / 1e9 / pcie_gbs if pcie_gbs else None}
            if detail:
                out[key]["stages"] = detail
        out["host_reader"]["threads"] = cores
        out["output_bytes"] = len(ref)
        out["speedup_over_host_reader"] = out["host_reader"]["stream_s"] / out["gpu_decoder"]["stream_s"]
        out["wall_speedup_over_host_reader"] = out["host_reader"]["wall_s_with_process_setup"] / out["gpu_decoder"]["wall_s_with_process_setup"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json configs[i]")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cigar-kernel", type=int, default=0, help="0 auto (default), 1 warp per record, 2 flat block scan, 3 streaming screen + thread per record")
    ap.add_argument("--reads-per-cta", type=int, default=0)
    ap.add_argument("--pipeline-parts", type=int, default=3, help="sub-batches in flight for the e2e measurement")
    ap.add_argument("--resident-batches", type=int, default=4, help="device-resident copies of the batch the timed steps go round robin over (1: one batch, L2 flushed before every step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true", help="the headline workload itself sharded round robin over the ranks (instead of one batch per rank)")
    ap.add_argument("--no-strong-c5", action="store_true", help="skip the strong_c5 block (N > 1)")
    ap.add_argument("--strong-scale", type=float, default=1.0, help="scale of configs[4] in the strong_c5 block")
    ap.add_argument("--no-e2e-bam", action="store_true", help="skip the e2e_bam block (BAM file -> event file through the CLI, N = 1)")
    ap.add_argument("--bam-scale", type=float, default=0.03, help="molecules of configs[1] in the e2e_bam BAM (0.03 -> ~31k records, ~580 MB)")
    ap.add_argument("--wc-input", action="store_true", help="write-combined pinned input views (A/B for the e2e figure)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of every step directly (default: repeated shapes replay a CUDA graph)")
    ap.add_argument("--no-overlap", action="store_true", help="run kernel 1 on the same stream as the SA branch")
    ap.add_argument("--fork-late", action="store_true", help="A/B: start the CIGAR path once kernel 0 is done (EXLR_OPT_OVERLAP = 2)")
    ap.add_argument("--k3-fold", action="store_true", help="A/B: kernel 3a's work inside kernel 3b for short-CIGAR batches (EXLR_OPT_K3_FOLD)")
    ap.add_argument("--k1-ctas", type=int, default=0, help="persistent CTAs of kernel 1 per SM (1..4)")
    ap.add_argument("--k1-waves", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    from excord_lr_b200 import api
    from excord_lr_b200.batch import ExlrParams

    torch.cuda.set_device(local_rank)
    D = Dist(world, local_rank)

    hb, c, _whole = make_workload(args.config, args.scale, rank, world, args.strong)
    p = ExlrParams.make(**c["params"])
    R, Cops, A = hb.n_reads, hb.n_ops, hb.n_sa_bytes

    ex_opts = [(api.EXLR_OPT_CIGAR_KERNEL, args.cigar_kernel), (api.EXLR_OPT_READS_PER_CTA, args.reads_per_cta)]
    if args.no_overlap:
        ex_opts.append((api.EXLR_OPT_OVERLAP, 0))
    if args.fork_late:
        ex_opts.append((api.EXLR_OPT_OVERLAP, 2))
    if args.k3_fold:
        ex_opts.append((api.EXLR_OPT_K3_FOLD, 1))
    if args.no_graph:
        ex_opts.append((api.EXLR_OPT_GRAPH, 0))
    if args.wc_input:
        ex_opts.append((api.EXLR_OPT_WC_INPUT, 1))
    if args.k1_ctas:
        ex_opts.append((api.EXLR_OPT_K1_CTAS_PER_SM, args.k1_ctas))
    if args.k1_waves:
        ex_opts.append((api.EXLR_OPT_K1_WAVES, args.k1_waves))
    ex = api.Extractor(p, hb.ref_names, local_rank)
    for k, v in ex_opts:
        ex.set_option(k, v)

    # ---------------- device-resident: value + roofline ----------------
    big = ex.batch_for(hb)
    big.upload()
    flush = Flusher(torch)
    # the timed steps run without the per-stage events (they cost a few microseconds of launch gap each); stage times and the
    # per-kernel durations are taken in extra steps below
    ex.set_option(api.EXLR_OPT_STAGE_TIMING, 0)
    sampler = ClockSampler(local_rank)
    sampler.start()
    wall0 = time.perf_counter()
    single_ms, launches, n_events, resident_step = measure_resident(big, flush, args.steps, args.warmup, D)
    total_ms, n_inflight, piped_ms, piped_copies = single_ms, 1, None, 0
    if args.resident_batches > 1:
        copies = [big]
        try:
            for _ in range(args.resident_batches - 1):
                b2 = ex.batch_for(hb); b2.upload(); copies.append(b2)
        except Exception:                                                  # (a config too large for that many copies: fewer)
            pass
        if len(copies) > 1:
            piped_ms, piped_launches, ne2 = measure_pipelined(copies, args.steps, args.warmup, D)
            assert ne2 == n_events
            piped_copies = len(copies)
            # a caller keeps as many batches in flight as serves its workload: long-CIGAR batches (HBM-bound from the first to the
            # last microsecond) gain nothing from a second step beside them and lose a little to the contention
            if max(D.reduce([piped_ms], "max")) < max(D.reduce([single_ms], "max")):
                total_ms, launches, n_inflight = piped_ms, piped_launches, piped_copies
        for b2 in copies[1:]:
            b2.free()
    wall_resident = time.perf_counter() - wall0
    counters = big.counters().as_dict()

    def diag(n):
        acc, k1, k1a = {}, [], []
        for i in range(3 + n):
            _r, t = resident_step()
            if i >= 3:
                k1.append(t.cigar_ms)
                k1a.append(t.screen_ms)
                for k, v in t.as_dict().items():
                    if k.endswith("_ms"):
                        acc[k] = acc.get(k, 0.0) + v / n
        return acc, k1, k1a
    ex.set_option(api.EXLR_OPT_STAGE_TIMING, 1)
    nd = min(args.steps, 20)
    stage_ms, k1_beside, _ = diag(nd)                                      # as configured (kernel 1 beside the SA branch)
    k1_overlapped_ms = float(np.mean(k1_beside)) if k1_beside else float("nan")
    # every kernel on its own: all on one stream, so no launch duration is stretched by the other branch
    if not args.no_overlap and not c["params"].get("split_only"):
        ex.set_option(api.EXLR_OPT_OVERLAP, 0)
        solo_stage, k1_ms, k1a_ms = diag(nd)
        ex.set_option(api.EXLR_OPT_OVERLAP, 2 if args.fork_late else 1)
    else:
        solo_stage, k1_ms, k1a_ms = dict(stage_ms), k1_beside, [0.0]
    ex.set_option(api.EXLR_OPT_STAGE_TIMING, 0)
    clocks = sampler.stop()
    big.free()

    # ---------------- PCIe ceiling: alone, and with every rank copying at once ----------------
    pcie_alone = pcie_probe(D, False)
    pcie_together = pcie_probe(D, True) if world > 1 else pcie_alone

    # ---------------- end to end through the public API, host buffers ----------------
    e2e_s, h2d, d2h, n_parts = measure_e2e(ex, hb, args.pipeline_parts, args.steps, args.warmup, n_events, D)

    # ---------------- reduce over ranks ----------------
    total_ms_max, e2e_ms_max, k1_ms_max, k1a_ms_max, single_ms_max, piped_ms_max = D.reduce([total_ms, e2e_s * 1e3, float(np.sum(k1_ms)), float(np.sum(k1a_ms)), single_ms, piped_ms if piped_ms is not None else 0.0], "max")
    R_all, C_all, E_all, pcie_sum, h2d_all = D.reduce([R, Cops, n_events, pcie_together, h2d], "sum")
    pcie_slowest, = D.reduce([pcie_together], "min")

    # ---------------- cpu baseline (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_c
        oracle_c.run(hb.slice(0, min(R, 1000)), p)          # load + warm the library outside the timed call
        t0 = time.perf_counter()
        want = oracle_c.run(hb, p)
        dt = time.perf_counter() - t0
        assert want.status == 0
        cpu = {"value": R / dt, "unit": "alignments/s", "cores": 1, "kind": "port",
               "sample": f"the whole workload once ({R} records, {Cops} CIGAR ops, {dt:.2f} s), single thread like the reference loop",
               "lines": int(len(want.events)), "matches_gpu_line_count": bool(len(want.events) == n_events)}

    bam_block = None
    if rank == 0 and world == 1 and not args.no_e2e_bam:
        ex.close()
        ex = None
        torch.cuda.synchronize()
        bam_block = e2e_bam(args, pcie_alone)

    strong = None
    if world > 1 and not args.no_strong_c5 and not args.strong:
        strong = strong_c5(args, ex_opts, D, rank, local_rank, world, load_peaks()[0])

    if rank == 0:
        peak, peak_src = load_peaks()
        traffic = load_traffic()
        # The CIGAR path is kernel 1a (streaming event screen: reads every op once and lists the 512-op steps that hold an event
        # candidate) followed by kernels 1b/1d on the records around those steps; event-dense batches get kernel 1 (flat block
        # scan of everything) instead.  The headline roofline entry is the kernel that moves the bytes: 1a reads 4 B/op of the
        # 4 B/op + 28 B/record the whole step reads.  Its peak is the measured COPY bandwidth (read + write); a read-mostly
        # stream can exceed that figure, so frac may come out slightly above 1 on long-record batches.
        screened = bool(k1a_ms) and float(np.mean(k1a_ms)) > 0
        path_bytes = 4.0 * Cops + 11.0 * R                                # 4 B/op + offset 8 + flag 2 + mapq 1 per record
        path_s = float(np.mean(k1_ms)) / 1e3 if k1_ms else float("nan")
        if screened:
            rk_name = "k1a_screen"
            k1_bytes = 4.0 * Cops                                         # every op read once (the list of flagged steps is a few KB)
            k1_avg_s = float(np.mean(k1a_ms)) / 1e3
        else:
            rk_name = "k1_flat" if args.cigar_kernel != 1 else "k1_warp"
            k1_bytes, k1_avg_s = path_bytes, path_s
        achieved = k1_bytes / k1_avg_s / 1e9 if k1_avg_s > 0 else 0.0
        kernels = kernel_table(hb, p, solo_stage, counters, n_events, peak, traffic) if screened or c["params"].get("split_only") else []
        dominant = max(kernels, key=lambda k: k["avg_ms"]) if kernels else None
        pipe_bytes = 24.0 * R + 4.0 * Cops + A + 4.0 * R + 48.0 * n_events
        ms_per_step = total_ms_max / args.steps
        value = R_all / (total_ms_max / 1e3) * args.steps
        e2e_value = R_all / (e2e_ms_max / 1e3) * args.steps
        h2d_rate = h2d * args.steps / (e2e_ms_max / 1e3) / 1e9             # this rank's bytes over the slowest rank's time
        line = {
            "metric": "alignments_per_sec", "value": value, "unit": "alignments/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "cigar_ops_per_sec": C_all / (total_ms_max / 1e3) * args.steps,
            "config": config_dict(c, hb),
            "single_batch": {"ms_per_step": single_ms_max / args.steps, "value": R_all / (single_ms_max / 1e3) * args.steps, "unit": "alignments/s",
                             "what": "ONE resident batch stepped alone: L2 flushed (256 MB read) before every step, first kernel start -> result header on the "
                                     "host by the library's own CUDA events (the figure earlier rounds reported as value)"},
            "pipelined": ({"ms_per_step": piped_ms_max / args.steps, "value": R_all / (piped_ms_max / 1e3) * args.steps, "unit": "alignments/s", "copies_in_flight": piped_copies,
                           "what": "steps round robin over device-resident copies of the batch, each on its own streams, a copy waited right before it is "
                                   "submitted again; device events around the K steps; the copies together exceed L2 several times, no flush kernel"}
                          if piped_ms is not None else None),
            "run": {"lines_per_gpu": n_events, "n_gpus": world, "resident_batches_in_flight": n_inflight,
                    "value_is": "pipelined" if n_inflight > 1 else "single_batch",
                    "launch": "direct launches" if args.no_graph else "the step's kernels replay as one CUDA graph (same shape every step); PDL edges and the two-stream fork/join are part of the graph",
                    "cigar_kernel": {0: "auto: streaming event screen (1a), then only the records around a candidate are scanned (1b: thread per short record; 1d: long records from 1a's per-step sums)",
                                     1: "warp per record", 2: "flat TMA-staged block scan of everything", 3: "the screened path (forced)"}[args.cigar_kernel],
                    "counters": counters},
            "e2e": {"value": e2e_value, "unit": "alignments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms_max / args.steps, "pipeline_parts": n_parts,
                    "h2d_gbs": h2d_rate, "h2d_gbs_all_gpus": h2d_all * args.steps / (e2e_ms_max / 1e3) / 1e9,
                    "pcie_h2d_peak_gbs": pcie_alone,
                    "pcie_h2d_peak_gbs_all_ranks_copying": pcie_together, "pcie_h2d_box_aggregate_gbs": pcie_sum,
                    "frac_of_pcie": h2d_rate / pcie_alone if pcie_alone else None,
                    "frac_of_box_aggregate": (h2d_all * args.steps / (e2e_ms_max / 1e3) / 1e9) / pcie_sum if pcie_sum else None,
                    # the e2e time is the slowest rank's: its own ceiling is what that rank's copies reach while all ranks copy
                    "pcie_h2d_slowest_rank_gbs_all_ranks_copying": pcie_slowest,
                    "frac_of_slowest_rank_ceiling": h2d_rate / pcie_slowest if pcie_slowest else None,
                    "cigar_ops_per_sec": C_all / (e2e_ms_max / 1e3) * args.steps,
                    "note": "the packer is outside the timed region (the pinned views are pre-filled); every step's inputs cross PCIe and every result is read back"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": rk_name, "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": k1_bytes, "avg_launch_ms": k1_avg_s * 1e3,
                         "measured": "live CUDA events in bench.py, same-stream steps (the kernel alone on the GPU)",
                         "cigar_path": {"kernels": "k1a_screen + k1b_claim + k1b_walk (+ k1d_long in long-record batches)" if screened else rk_name,
                                        "algorithmic_bytes": path_bytes, "ms": path_s * 1e3,
                                        "achieved": path_bytes / path_s / 1e9 if path_s > 0 else 0.0,
                                        "frac": (path_bytes / path_s / 1e9 / peak) if path_s > 0 else 0.0,
                                        "ms_beside_sa_branch": k1_overlapped_ms},
                         "traffic": (traffic.get(rk_name) or {}).get("dram_bytes_per_launch"),
                         "kernels": kernels,
                         "time_dominant_kernel": dominant["name"] if dominant else None,
                         "time_dominant_frac": dominant["frac"] if dominant else None,
                         "pipeline_achieved_gbs": pipe_bytes / (ms_per_step / 1e3) / 1e9,
                         "pipeline_frac": pipe_bytes / (ms_per_step / 1e3) / 1e9 / peak},
            "stage_ms": stage_ms, "stage_ms_same_stream": solo_stage,
            "clocks": clocks,
            "wall_s_resident_loop": wall_resident,
        }
        if cpu:
            line["cpu_baseline"] = cpu
        if strong:
            line["strong_c5"] = strong
        if bam_block:
            line["e2e_bam"] = bam_block
        print(json.dumps(line), flush=True)

    if ex is not None:
        ex.close()
    D.close()


if __name__ == "__main__":
    main()
