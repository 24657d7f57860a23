#!/usr/bin/env python
"""bench.py — polished Mbp/s of the B200 polish path on BASELINE.json's configs[1]
(synthetic 10 Mbp haploid contig, 30x HiFi, k21+k31), one contig per GPU (contigs are independent: weak scaling,
no collective on the data path).

    python bench.py --gpus N --steps K --warmup W             # our arm
    python bench.py --impl reference --steps K --warmup W     # the reference's CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  `value` = polished Mbp/s with the inputs already resident in HBM when the
timed region starts; `e2e` = the same metric through np2_polish_contig with HOST buffers (H2D + D2H inside the
timed region).  `roofline` describes the dominant kernel of the step, `yak_lookup` the named K5 kernel at a
table size far above L2, `cpu_baseline` the oracle port on a bounded sample of the same workload.
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

import numpy as np  # noqa: E402

WORKLOAD = "synthetic 10 Mbp haploid contig, 30x HiFi (N(15k,2k), 0.2% err), asm err 2e-5/bp, k21+k31"
WORKLOAD_DIPLOID = ("synthetic %.0f Mbp diploid contig (%.1f%% het), 30x HiFi from both haplotypes, asm err 2e-5/bp, k21+k31 "
                    "(one contig of configs[2]; exercises phasing)")
KS = (21, 31)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_workload(seed, length, threads, het=0.0):
    from nextpolish2_b200 import synth
    t0 = time.time()
    A = synth.genome(seed, length)
    c = synth.make_contig(seed + 1, A, depth=30.0, asm_err=2e-5, het=het, read_err=0.002, threads=threads)
    tabs = {k: synth.make_table(seed + 2, k, [c["hap1"]] + ([c["hap2"]] if het > 0 else [])) for k in KS}
    log("[bench] workload seed %d: %d bp, %d reads, %.1f MB of BAM records, tables %s  (%.1fs)" % (
        seed, length, c["n_reads"], len(c["bam"]) / 1e6, {k: len(v[0]) for k, v in tabs.items()}, time.time() - t0))
    return A, c, tabs


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# the kernel behind each single-kernel stage as ncu lists it (defaults: NP2_PACK_BATCH=2, NP2_PILE_BATCH=2)
KERNEL_OF = {"pack_columns": "k_pack_columns_batched<2>", "pileup_emit": "k_pileup_emit<2>"}
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel, from the committed ncu capture of THIS workload
# (profiles/r01ab_ncu_variants.txt, 10 Mbp / 30x launch; averages over the launches), bytes.  None = not captured.
NCU_TRAFFIC = {"pack_columns": 384700000, "pileup_emit": 273600000}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


# Algorithmic bytes of the single-kernel stages (DESIGN.md "Kernels"): what the kernel must move through HBM once.
def stage_bytes(stage, st):
    cols, L = st["alignment_columns"], st["L"]
    return {
        # 4-bit SEQ in (0.5 B/column), packed nibble out (0.5), one 10-byte checkpoint out and one 2-byte op index in
        # per 32 columns
        "pack_columns": cols * (0.5 + 0.5 + 10 / 32 + 2 / 32),
        # packed columns in (0.5) + checkpoints in (10/32); the packed reference (0.5 B/bp) is shared by the ~30 reads
        # over a position and counted once
        "pileup_emit": cols * (0.5 + 10 / 32) + L * 0.5 + st.get("records", 0) * 12,
    }.get(stage)


# bytes the whole pipeline must move per polished bp, D = depth, q = yak probes per bp
def pipeline_bytes_per_bp(depth, q):
    return 1.5 * depth + 111 + 42 * q


def run_ours(args):
    import torch
    import torch.distributed as dist
    import nextpolish2_b200 as np2
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    threads = max(1, (os.cpu_count() or 8) // max(world, 1))
    A, c, tabs = make_workload(20260002 + 1000 * rank, args.length, min(threads, 16), args.het)
    workload = WORKLOAD if args.het == 0 else WORKLOAD_DIPLOID % (args.length / 1e6, args.het * 100)
    ctx = np2.Context(local)
    # the box's cores are shared by the ranks and by the contigs each rank keeps in flight
    from nextpolish2_b200.api import set_host_threads
    set_host_threads(args.host_threads or
                     max(2, min(16, (os.cpu_count() or 8) // max(world, 1) // max(1, min(args.e2e_inflight, 2)) * 2)))
    tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in KS]
    opts = np2.Opts()  # reference defaults; the 10 Mbp contig is above -L 1000000
    bam_pinned = torch.from_numpy(c["bam"]) if args.pageable else torch.from_numpy(c["bam"]).pin_memory()
    contig_pinned = torch.from_numpy(A.copy()).pin_memory()
    bam_np, contig_np = bam_pinned.numpy(), contig_pinned.numpy()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm: inputs uploaded once, the timed step is np2_job_run
    job = np2.Job(ctx, contig_np, bam_np, tables, opts).upload()
    for _ in range(args.warmup):
        job.run(-1)
    barrier()
    sampler = ClockSampler(local)
    step_ms, stage_acc, launches = [], {}, 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        job.run(-1)
        tm = job.timings()
        step_ms.append(tm["total"][0])
        for k, v in tm.items():
            stage_acc.setdefault(k, []).append(v[0])
        stage_launches = {k: v[1] for k, v in tm.items()}
        launches += job.traffic()["kernel_launches"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    traffic = job.traffic()
    _, _, gbase = job.bases()
    dev_time = sum(step_ms) / 1e3  # CUDA events on the library's stream around each whole step (includes host phases)
    ident = bytes(gbase) == bytes(c["hap1"])
    job.destroy()

    # ---- end to end: host buffers in, consensus out, every step.  `--e2e-inflight` contigs are in flight at once
    # (one host thread + one context + one stream each, tables shared: the CLI's double buffering), so the PCIe
    # upload and host-side record parsing of one contig overlap the kernels of another.
    e2e_parts = {"create_parse": 0.0, "upload": 0.0, "run": 0.0, "result": 0.0, "n": 0}
    e2e_parts_inflight = {"create_parse": 0.0, "upload": 0.0, "run": 0.0, "result": 0.0, "n": 0}

    def e2e_step(cx, acc=None):
        t0 = time.perf_counter()
        j = np2.Job(cx, contig_np, bam_np, tables, opts)
        t1 = time.perf_counter()
        j.upload()
        t2 = time.perf_counter()
        j.run(-1)
        t3 = time.perf_counter()
        first, last, base = j.bases(copy=False)  # the FASTA record (header span + bases) in host memory
        assert len(base) == len(gbase) and base[-1] == gbase[-1] and base[len(base) // 2] == gbase[len(base) // 2]
        tr = j.traffic()
        tm_up = {k: v[0] for k, v in j.timings().items() if k.startswith("upload:")} if acc is not None else {}
        j.destroy()
        t4 = time.perf_counter()
        if acc is not None:
            for k, v in zip(("create_parse", "upload", "run", "result"), (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                acc[k] += v * 1e3
            for k, v in tm_up.items():
                acc[k] = acc.get(k, 0.0) + v
            acc["n"] += 1
        return tr

    def e2e_run(n_inflight):
        ctxs = [ctx] + [np2.Context(local) for _ in range(n_inflight - 1)]
        for cx in ctxs:  # warm every context's pools
            for _ in range(max(1, args.warmup // n_inflight)):
                e2e_step(cx)
        tr_box, errs = [None], []
        share = [args.steps // n_inflight + (1 if w < args.steps % n_inflight else 0) for w in range(n_inflight)]

        def work(w):
            try:
                for _ in range(share[w]):
                    tr_box[0] = e2e_step(ctxs[w], e2e_parts if n_inflight == 1 else e2e_parts_inflight)
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        torch.cuda.synchronize()
        th = [threading.Thread(target=work, args=(w,)) for w in range(n_inflight)]
        t1 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for cx in ctxs[1:]:
            cx.close()
        if errs:
            raise errs[0]
        return t2 - t1, tr_box[0]

    e2e_serial_time, tr = e2e_run(1)
    e2e_time = e2e_serial_time
    if args.e2e_inflight > 1:
        e2e_time, tr = e2e_run(args.e2e_inflight)
    clocks = sampler.stop()  # sampled over both timed regions (device-resident steps and end-to-end steps)

    t_dev = torch.tensor([dev_time, e2e_time, wall, e2e_serial_time], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_time, e2e_time, wall, e2e_serial_time = [float(x) for x in t_dev.tolist()]
    mbp_total = world * args.length * args.steps / 1e6

    line = None
    if rank == 0:
        peak, peak_kind = peaks()
        stages = {k: float(np.mean(v)) for k, v in stage_acc.items()}
        st = {"alignment_columns": traffic["alignment_columns"], "L": args.length, "records": 0}
        kern = {k: v for k, v in stages.items() if stage_bytes(k, st)}
        dom = max(kern, key=kern.get) if kern else None
        roofline = None
        if dom:
            # single-kernel stages; the stage time is summed over the launches of the step
            n_launch = max(1, int(round(stage_launches.get(dom, 1))))
            per_launch_ms = stages[dom] / n_launch
            achieved = stage_bytes(dom, st) / (per_launch_ms * 1e-3) / 1e9
            roofline = {"kernel": KERNEL_OF.get(dom, "k_" + dom), "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                        "frac": round(achieved / peak, 4), "traffic": NCU_TRAFFIC.get(dom), "peak_source": peak_kind,
                        "launch_ms": round(per_launch_ms, 4), "launches_per_step": n_launch,
                        "share_of_step": round(stages[dom] / stages["total"], 4),
                        "algorithmic_bytes_per_launch": int(stage_bytes(dom, st)),
                        "note": "largest of the streaming kernels with a defined byte count (pack_columns, "
                                "pileup_emit); the step is ~50 short kernels + host phases, none above 10% of it"}
        others = []
        for kname in kern:
            if kname == dom:
                continue
            nl = max(1, int(round(stage_launches.get(kname, 1))))
            ms1 = stages[kname] / nl
            ach = stage_bytes(kname, st) / (ms1 * 1e-3) / 1e9
            others.append({"kernel": KERNEL_OF.get(kname, "k_" + kname), "achieved": round(ach, 1), "frac": round(ach / peak, 4),
                           "launch_ms": round(ms1, 4), "traffic": NCU_TRAFFIC.get(kname)})
        if roofline is not None:
            roofline["other_streaming_kernels"] = others
        q = traffic["probes"] / float(args.length)
        pipe_bytes = pipeline_bytes_per_bp(30, q) * args.length
        pipeline = {"bytes_per_bp": round(pipeline_bytes_per_bp(30, q), 1), "probes_per_bp": round(q, 3),
                    "achieved_GBps": round(pipe_bytes / (stages["total"] * 1e-3) / 1e9, 1),
                    "frac_of_hbm_peak": round(pipe_bytes / (stages["total"] * 1e-3) / 1e9 / peak, 4)}
        line = {
            "metric": "polished Mbp/s", "value": round(mbp_total / dev_time, 3), "unit": "Mbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_time / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer",
            "data": "synthetic",
            "config": {"workload": workload, "contig_bp": args.length, "contigs_per_gpu": 1, "depth": 30,
                       "tables": "k21+k31 synthesised from the truth haplotype(s)", "partition": "one contig per GPU, no collective",
                       "l2": "inputs (%.0f MB of BAM records per GPU) are larger than the 126 MB L2" % (len(c["bam"]) / 1e6)},
            "e2e": {"value": round(mbp_total / e2e_time, 3), "unit": "Mbp/s", "h2d_bytes_per_step": tr["h2d_bytes"],
                    "d2h_bytes_per_step": tr["d2h_bytes"], "contigs_in_flight": args.e2e_inflight,
                    "one_at_a_time": round(mbp_total / e2e_serial_time, 3),
                    "one_at_a_time_ms": {k: round(v / max(1, e2e_parts["n"]), 3) for k, v in e2e_parts.items() if k != "n"},
                    "in_flight_ms_per_contig_per_thread": {k: round(v / max(1, e2e_parts_inflight["n"]), 3)
                                                           for k, v in e2e_parts_inflight.items()
                                                           if k in ("create_parse", "upload", "run", "result")}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "pipeline_roofline": pipeline,
            "stages_ms": {k: round(v, 4) for k, v in stages.items()},
            "identical_to_truth_haplotype": bool(ident),
            "wall_ms_per_step": round(wall / args.steps * 1e3, 3),
        }
        if not args.no_yak_bench:
            line["yak_lookup"] = yak_bench(ctx, np2, torch, peak)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(args, steps=args.cpu_steps)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def yak_bench(ctx, np2, torch, peak):
    """K5 at scale: 2^26 keys (~0.9 GB table >> L2), 2^26 device-resident queries, half present / half absent."""
    n = 1 << 26
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    keys = torch.randint(0, 2**62, (n,), dtype=torch.int64, device="cuda", generator=g)
    keys = torch.unique(keys)
    nk = keys.numel()
    cnt = ((keys >> 13) % 1023 + 1).to(torch.int16)
    tab = np2.Table.from_arrays(ctx, 31, keys.cpu().numpy().view(np.uint64), cnt.cpu().numpy().view(np.uint16))
    perm = torch.randperm(nk, device="cuda", generator=g)[: n // 2]
    q = torch.cat([keys[perm], torch.randint(0, 2**62, (n // 2,), dtype=torch.int64, device="cuda", generator=g)])
    q = q[torch.randperm(q.numel(), device="cuda", generator=g)].contiguous()
    out = torch.empty(q.numel(), dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    tab.lookup_device(q.data_ptr(), q.numel(), out.data_ptr(), 5, repeat=3)
    ms = tab.lookup_device(q.data_ptr(), q.numel(), out.data_ptr(), 5, repeat=10)
    # size-independent check: present keys answer their count (>= 5 filter), absent answer 0
    exp = torch.where(cnt[perm] >= 5, cnt[perm], torch.zeros_like(cnt[perm]))
    ok = bool((tab.lookup(keys[perm].cpu().numpy().view(np.uint64), 5) == exp.cpu().numpy().view(np.uint16)).all())
    probes = q.numel()
    from nextpolish2_b200.api import bench_gather32
    g_ms = bench_gather32(ctx, tab.device_bytes, probes, repeat=10)
    gather_gbs = probes * 32 / g_ms / 1e6
    res = {"probes": probes, "table_keys": nk, "table_bytes": tab.device_bytes, "ms": round(ms, 4),
           "gprobes_per_s": round(probes / ms / 1e6, 3),
           "sector_GBps": round(probes * 32 / ms / 1e6, 1), "algorithmic_GBps": round(probes * 42 / ms / 1e6, 1),
           "frac_of_stream_peak": round(probes * 42 / ms / 1e6 / peak, 4),
           "random_gather_peak_GBps": round(gather_gbs, 1), "frac_of_random_gather_peak": round(probes * 32 / ms / 1e6 / gather_gbs, 4),
           "correct": ok}
    tab.free()
    return res


def cpu_sample(args, steps=1, warmup=0):
    """The reference's CPU algorithm (oracle port, in-memory tables) on a bounded sample of the same workload:
    `cores` contigs of 1 Mbp each, one worker thread per contig (the reference's unit of parallelism)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from nextpolish2_b200 import synth
    cores = min(os.cpu_count() or 1, 32)
    clen = args.cpu_contig
    G = synth.genome(20260002, clen * cores)
    data = []
    for i in range(cores):
        A = G[i * clen:(i + 1) * clen].copy()
        data.append((A, synth.make_contig(20260100 + i, A, depth=30.0, asm_err=2e-5, het=args.het, read_err=0.002, threads=4)))
    haps = [d[1]["hap1"] for d in data] + ([d[1]["hap2"] for d in data] if args.het > 0 else [])
    tabs = [O.Table.from_arrays(k, *synth.make_table(20260003, k, haps)) for k in KS]
    opts = O.Opts(min_ctg_len=0)
    vals = []
    for it in range(warmup + steps):
        res = [None] * cores

        def work(i):
            res[i] = O.Job(data[i][0], data[i][1]["bam"], tabs, opts, dump_iter=-1)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        [t.start() for t in th]
        [t.join() for t in th]
        dt = time.perf_counter() - t0
        assert args.het > 0 or all(bytes(res[i].consensus()[1]) == bytes(haps[i]) for i in range(cores))
        if it >= warmup:
            vals.append(cores * clen / 1e6 / dt)
    return {"value": round(float(np.mean(vals)), 4), "unit": "Mbp/s", "cores": cores, "kind": "port",
            "sample": "%d contigs x %d bp of the same synthetic workload (30x), one oracle thread per contig, "
                      "tables in memory (kinder to the CPU than the reference's per-pass .yak file streaming); "
                      "restatement of the reference algorithm, not the Rust binary (no cargo in this image)" % (cores, clen),
            "seconds": round(len(vals) * cores * clen / 1e6 / float(np.mean(vals)), 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    b = cpu_sample(args, steps=max(1, args.steps), warmup=min(args.warmup, 2))
    line = {"impl": "reference", "metric": "polished Mbp/s", "value": b["value"], "unit": "Mbp/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", args.gpus)), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(b["seconds"] * 1e3 / max(1, args.steps), 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/u64 integer", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": b["sample"]},
            "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": b["value"], "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--length", type=int, default=10_000_000, help="contig length per GPU (configs[1] = 10 Mbp)")
    ap.add_argument("--cpu-contig", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=10, help="passes over the CPU sample for cpu_baseline (~1 s each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--het", type=float, default=0.0, help="heterozygosity of the synthetic contig (configs[2]: 0.01)")
    ap.add_argument("--no-yak-bench", action="store_true")
    ap.add_argument("--pageable", action="store_true", help="record buffer in pageable memory (host compaction path)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per library call (0 = cores / ranks / ~in flight)")
    ap.add_argument("--e2e-inflight", type=int, default=3, help="contigs in flight per GPU in the end-to-end arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
