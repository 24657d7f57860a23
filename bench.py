#!/usr/bin/env python
"""bench.py — polished Mbp/s of the B200 polish path (BASELINE.json: "polished Mbp/sec at 1/2/4/8 B200 vs reference
CPU; identical FASTA output").

    python bench.py --gpus N --steps K --warmup W              # our arm (headline: configs[1], one contig per GPU)
    python bench.py --config 2|3 [--full]                      # another config as the headline workload
    python bench.py --impl reference --steps K --warmup W      # the reference's CPU algorithm (oracle port)

One JSON line on stdout (rank 0).  A "step" is one pass of the polish path over the workload's contigs.
  value     polished Mbp/s with the inputs already resident in HBM when the timed region starts (np2_job_run)
  e2e       the same metric through np2_job_create/upload/run with HOST buffers (H2D + D2H inside the timed region)
  verify    SHA-256 of the FASTA the GPU produced vs the FASTA of the CPU oracle on the same inputs
  configs   (N = 1) the other BASELINE.json configs at a bounded size, each with its own value / e2e / verify:
            "2" = 10 Mbp diploid contigs (1 % het, phasing), "3" = tandem-repeat contig, 40x, k21+k31+k51
  strong_scaling   24 diploid contigs with configs[4]'s length distribution (scaled), LPT-partitioned over the N ranks,
            tables staged once and broadcast over NVLink, FASTA digest in input order (must be the same at every N)
  roofline  the largest single-kernel stage of the step against the measured HBM peak; yak_lookup = K5 at scale;
  cpu_baseline     the oracle port on a bounded sample of the same workload.
Full-size runs (`--config 2 --full`, `--config 3 --full`) are kept under profiles/.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# BASELINE.json configs (sizes: SURVEY.md 8d).  "bounded" = what the default run uses when the config is an extra.
CONFIGS = {
    1: {"label": "configs[1]", "contigs": 1, "length": 10_000_000, "het": 0.0, "tandem": 0.0, "depth": 30.0, "ks": (21, 31),
        "text": "synthetic 10 Mbp haploid contig, 30x HiFi (N(15k,2k), 0.2% err), asm err 2e-5/bp, k21+k31"},
    2: {"label": "configs[2]", "contigs": 10, "length": 10_000_000, "het": 0.01, "tandem": 0.0, "depth": 30.0, "ks": (21, 31),
        "bounded": {"contigs": 2},
        "text": "synthetic diploid (1%% het: 0.8%% SNV + 0.2%% indel), %d x %.0f Mbp contigs, 30x HiFi from both haplotypes, "
                "asm err 2e-5/bp, k21+k31 (exercises phasing)"},
    3: {"label": "configs[3]", "contigs": 1, "length": 250_000_000, "het": 0.0, "tandem": 0.05, "depth": 40.0, "ks": (21, 31, 51),
        "bounded": {"length": 20_000_000},
        "text": "synthetic chr1-scale contig, %d x %.0f Mbp, 5%% tandem-repeat blocks (unit 2-60 bp), 40x HiFi, "
                "k21+k31+k51 (k51 = bit-plane hash)"},
}
SEED0 = 20260000  # seed = 20260000 + config number


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def config_text(cfg):
    return cfg["text"] % (cfg["contigs"], cfg["length"] / 1e6) if "%d x" in cfg["text"] else cfg["text"]


def make_workload(cfg, seed, threads):
    """[(name, contig, hap1, hap2, bam)], {k: (hashes, counts)} of one config (deterministic)."""
    from nextpolish2_b200 import synth
    t0 = time.time()
    contigs = []
    for i in range(cfg["contigs"]):
        A = synth.genome(seed + 10 * i, cfg["length"], tandem_frac=cfg["tandem"])
        c = synth.make_contig(seed + 10 * i + 1, A, depth=cfg["depth"], asm_err=2e-5, het=cfg["het"], read_err=0.002,
                              threads=threads)
        contigs.append({"name": "ctg%03d" % i, "contig": A, "hap1": c["hap1"], "hap2": c["hap2"], "bam": c["bam"],
                        "n_reads": c["n_reads"]})
    t1 = time.time()
    haps = [c["hap1"] for c in contigs] + [c["hap2"] for c in contigs if len(c["hap2"])]
    tabs = {k: synth.make_table_mt(seed + 2, k, haps, threads=threads) for k in cfg["ks"]}
    log("[bench] %s seed %d: %d x %d bp, %d reads, %.1f MB of BAM records (%.1fs), tables %s (%.1fs)" % (
        cfg["label"], seed, cfg["contigs"], cfg["length"], sum(c["n_reads"] for c in contigs),
        sum(len(c["bam"]) for c in contigs) / 1e6, t1 - t0, {k: len(v[0]) for k, v in tabs.items()}, time.time() - t1))
    return contigs, tabs


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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------------------------
# Algorithmic bytes of the single-kernel stages: what the kernel must move through HBM once.
# name -> (kernel as ncu lists it, bytes(st)); st = {"cols", "L", "records", "groups", "N"}
STAGE_MODELS = {
    # 4-bit SEQ in (0.5 B/column), packed nibble out (0.5), one 10-byte checkpoint out + one 2-byte op index in per 32 columns
    # (+ one "not all reference" bit out per block and the packed reference, 0.5 B/bp, once: below 2 % of the rest)
    "pack_columns": ("k_pack_columns_batched<2>", lambda st: st["cols"] * (0.5 + 0.5 + 10 / 32 + 2 / 32)),
    # K2: per position coverage 4 + code 1 in, entry offset 4 + count 2 + reference count 4 + flag 1 + emit count 4 out;
    # one bit per 32-column block in; per not-all-reference block (<= one per record) 16 B of columns + 10 B checkpoint in;
    # 16 B per Msa entry out
    "pileup_stripe": ("k_pileup_stripe<1024, true>",
                      lambda st: st["L"] * 20 + st["cols"] / 256 + st["records"] * 26 + st["groups"] * 16),
    # per multi-entry position: offset 4 + count 2 + coverage 4 + flag 1 in, reference score / besti 12 out; per entry 16 in, 12 out
    "dp_runs": ("k_dp_runs (+ k_dp_runs_long)", lambda st: st["groups"] * (11 + 12) + st["groups"] * 28),
}
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of configs[1] (profiles/), bytes
NCU_TRAFFIC = {"pack_columns": 384700000}


# bytes the whole pipeline must move per polished bp, D = depth, q = yak probes per bp
def pipeline_bytes_per_bp(depth, q):
    return 1.5 * depth + 111 + 42 * q


def fasta_record(name, first, last, base):
    """display_consensusbase_vec (main.rs:627-643): header with the first/last consensus position, one line of bases."""
    return (">%s start:%d end:%d\n" % (name, first, last)).encode() + bytes(base) + b"\n"


def oracle_fasta(contigs, tabs, ks, opts_kw, threads, stream_scan=False):
    """FASTA records of the CPU oracle for every contig (one oracle thread per contig, `threads` at a time)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    otabs = [O.Table.from_arrays(k, *tabs[k]) for k in ks]
    for t in otabs:
        t.set_stream_scan(stream_scan)
    out, dropped, secs = [None] * len(contigs), [0] * len(contigs), [0.0] * len(contigs)
    nxt, lock = [0], threading.Lock()

    def work():
        while True:
            with lock:
                i = nxt[0]
                nxt[0] += 1
            if i >= len(contigs):
                return
            c = contigs[i]
            j = O.Job(c["contig"], c["bam"], otabs, O.Opts(**opts_kw), dump_iter=-1)
            pos, base = j.consensus()
            out[i] = fasta_record(c["name"], int(pos[0]), int(pos[-1]), base) if len(base) else b""
            dropped[i] = len(j.dropped())
            secs[i] = j.seconds
    th = [threading.Thread(target=work) for _ in range(max(1, min(threads, len(contigs))))]
    [t.start() for t in th]
    [t.join() for t in th]
    return out, dropped, secs


class Measured:
    pass


def measure(np2, torch, ctx, local, contigs, tables, cfg, args, opts, steps, warmup, barrier, inflight):
    """Device-resident arm + end-to-end arm over `contigs`; returns a Measured with times (seconds for `steps` steps)."""
    from nextpolish2_b200.api import set_stage_timing
    m = Measured()
    pinned = []
    for c in contigs:
        bam_t = torch.from_numpy(c["bam"]) if args.pageable else torch.from_numpy(c["bam"]).pin_memory()
        ctg_t = torch.from_numpy(c["contig"].copy()).pin_memory()
        pinned.append((ctg_t, bam_t, ctg_t.numpy(), bam_t.numpy()))
    m.pinned = pinned
    total_bp = sum(len(c["contig"]) for c in contigs)

    # ---- device-resident arm: inputs uploaded once, the timed step is np2_job_run over every contig
    jobs = [np2.Job(ctx, p[2], p[3], tables, opts).upload() for p in pinned]
    for _ in range(warmup):
        for j in jobs:
            j.run(-1)
    barrier()
    step_ms, stage_acc, launches, stage_launches = [], {}, 0, {}
    t0 = time.perf_counter()
    for _ in range(steps):
        tot, acc = 0.0, {}
        for j in jobs:
            j.run(-1)
            tm = j.timings()
            tot += tm["total"][0]
            for k, v in tm.items():
                acc[k] = acc.get(k, 0.0) + v[0]
                stage_launches[k] = v[1]
            launches += j.traffic()["kernel_launches"]
        step_ms.append(tot)
        for k, v in acc.items():
            stage_acc.setdefault(k, []).append(v)
    torch.cuda.synchronize()
    m.wall = time.perf_counter() - t0
    barrier()
    m.traffic = {k: sum(j.traffic()[k] for j in jobs) for k in jobs[0].traffic()}
    m.stats = {}
    for k in jobs[0].stats() if hasattr(jobs[0], "stats") else {}:
        m.stats[k] = sum(j.stats()[k] for j in jobs)
    m.records, m.dropped = [], []
    for c, j in zip(contigs, jobs):
        first, last, base = j.bases()
        m.records.append(fasta_record(c["name"], first, last, base))
        m.dropped.append(len(j.dropped()))
    m.dev_time = sum(step_ms) / 1e3  # CUDA events on the library's stream around each whole run (includes host phases)
    m.stages = {k: float(np.mean(v)) for k, v in stage_acc.items()}
    m.stage_launches = stage_launches
    m.launches = launches
    for j in jobs:
        j.destroy()

    # ---- end to end: host buffers in, FASTA record out, every step.  `inflight` contigs are in flight at once
    # (one host thread + one context + one stream each, tables shared: the CLI's scheme), so the PCIe upload and
    # host-side record parsing of one contig overlap the kernels of another.
    parts1 = {"create_parse": 0.0, "upload": 0.0, "run": 0.0, "result": 0.0, "n": 0}
    partsN = {"create_parse": 0.0, "upload": 0.0, "run": 0.0, "result": 0.0, "n": 0}
    expect = [len(r) for r in m.records]

    def e2e_begin(cx, ci):
        """np2_job_create: record parse on the host, then everything the upload needs is ENQUEUED (K0's gather of the SEQ
        fields over PCIe, the per-read arrays on the copy stream) and the call returns."""
        t0 = time.perf_counter()
        j = np2.Job(cx, pinned[ci][2], pinned[ci][3], tables, opts)
        return j, ci, time.perf_counter() - t0

    def e2e_finish(started, acc=None):
        j, ci, t_create = started
        c = contigs[ci]
        t1 = time.perf_counter()
        j.upload()
        t2 = time.perf_counter()
        j.run(-1)
        t3 = time.perf_counter()
        first, last, base = j.bases(copy=False)  # the FASTA record (header span + bases) in host memory
        rec_len = len(">%s start:%d end:%d\n" % (c["name"], first, last)) + len(base) + 1
        assert rec_len == expect[ci], "end-to-end FASTA record differs from the device-resident run"
        tr = j.traffic()
        tm = j.timings() if acc is not None else {}
        st = j.stats() if acc is not None else {}
        j.destroy()
        t4 = time.perf_counter()
        if acc is not None:
            for k, v in zip(("create_parse", "upload", "run", "result"), (t_create, t2 - t1, t3 - t2, t4 - t3)):
                acc[k] += v * 1e3
            acc["n"] += 1
            sa = acc.setdefault("stages", {})
            for k, v in tm.items():
                sa[k] = sa.get(k, 0.0) + v[0]
            acc["repeated_passes"] = acc.get("repeated_passes", 0) + st.get("repeated_passes", 0)
        return tr

    def e2e_one(cx, ci, acc=None):
        return e2e_finish(e2e_begin(cx, ci), acc)

    def e2e_run(n_inflight):
        # every worker thread owns `depth` contexts: while it runs contig i on one, contig i + 1 is already parsed and its
        # upload is under way on the other (np2_job_create only enqueues), so the link does not idle while the worker
        # is inside np2_job_run and no extra host thread competes for the cores
        depth = 2 if (args.e2e_prefetch and n_inflight > 1) else 1
        # every context grows its own device pool to what its largest contig needs (np2_job_create: ~4 x the record bytes
        # + 100 B per bp): bound the contexts by the memory that is free now
        want = max(len(c["bam"]) * 4 + len(c["contig"]) * 100 + (64 << 20) for c in contigs)
        fit = max(1, int(torch.cuda.mem_get_info()[0] * 0.8 // want))
        if n_inflight * depth > fit:
            depth = 2 if (depth == 2 and fit >= 2) else 1
            n_inflight = max(1, fit // depth)
        # stage timers (two event records per stage) only while one contig is processed at a time
        set_stage_timing(n_inflight == 1 or args.e2e_stage_timers)
        ctxs = [ctx] + [np2.Context(local) for _ in range(n_inflight * depth - 1)]
        for cx in ctxs:  # warm every context: pools, page-locked buffers, and the sizes the speculative passes start from
            for _ in range(max(2, warmup // n_inflight)):
                for ci in range(len(contigs)):
                    e2e_one(cx, ci)
        work_items = [ci for _ in range(steps) for ci in range(len(contigs))]
        nxt, lock, errs, tr_sum = [0], threading.Lock(), [], {"h2d_bytes": 0, "d2h_bytes": 0}

        def take():
            with lock:
                i = nxt[0]
                nxt[0] += 1
            return i if i < len(work_items) else None

        def work(w):
            acc = parts1 if n_inflight == 1 else partsN
            try:
                mine, turn = ctxs[w * depth:(w + 1) * depth], 0
                i = take()
                cur = e2e_begin(mine[0], work_items[i]) if i is not None else None
                while cur is not None:
                    nx = None
                    if depth > 1:
                        i = take()
                        if i is not None:
                            turn ^= 1
                            nx = e2e_begin(mine[turn], work_items[i])
                    tr = e2e_finish(cur, acc)
                    with lock:
                        tr_sum["h2d_bytes"] += tr["h2d_bytes"]
                        tr_sum["d2h_bytes"] += tr["d2h_bytes"]
                    if depth == 1:
                        i = take()
                        nx = e2e_begin(mine[0], work_items[i]) if i is not None else None
                    cur = nx
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
        set_stage_timing(True)
        if errs:
            raise errs[0]
        return t2 - t1, {k: v // steps for k, v in tr_sum.items()}

    barrier()
    m.e2e_serial_time, m.e2e_traffic = e2e_run(1)
    m.e2e_time = m.e2e_serial_time
    if inflight > 1:
        barrier()
        m.e2e_time, m.e2e_traffic = e2e_run(inflight)
    m.parts1, m.partsN = parts1, partsN
    m.total_bp = total_bp
    return m


def verify(contigs, tabs, cfg, records, dropped_gpu, opts_kw, threads):
    t0 = time.time()
    orec, odrop, osec = oracle_fasta(contigs, tabs, cfg["ks"], opts_kw, threads)
    g = hashlib.sha256(b"".join(records)).hexdigest()
    o = hashlib.sha256(b"".join(orec)).hexdigest()
    truth = [bytes(r.split(b"\n", 2)[1]) == bytes(c["hap1"]) for r, c in zip(records, contigs)]
    return {"fasta_sha256_gpu": g, "fasta_sha256_oracle": o, "identical": g == o, "contigs_checked": len(contigs),
            "fasta_bytes": sum(len(r) for r in records), "reads_dropped_by_phasing_gpu": int(sum(dropped_gpu)),
            "reads_dropped_by_phasing_oracle": int(sum(odrop)), "identical_to_truth_haplotype": bool(all(truth)),
            "oracle_cpu_seconds": round(float(sum(osec)), 1), "oracle_wall_seconds": round(time.time() - t0, 1)}


def summarise(m, cfg, steps, world, peak):
    """value / e2e / pipeline roofline of one measured config (rank-local times; the caller reduces over ranks)."""
    q = m.traffic["probes"] / float(m.total_bp)
    bpb = pipeline_bytes_per_bp(cfg["depth"], q)
    step_ms = m.dev_time / steps * 1e3
    return {
        "workload": config_text(cfg), "contigs": cfg["contigs"], "contig_bp": cfg["length"],
        "value": round(world * m.total_bp * steps / 1e6 / m.dev_time, 3), "unit": "Mbp/s",
        "ms_per_step": round(step_ms, 3), "ms_per_10Mbp": round(step_ms * 1e7 / m.total_bp, 3),
        "e2e": {"value": round(world * m.total_bp * steps / 1e6 / m.e2e_time, 3), "unit": "Mbp/s",
                "h2d_bytes_per_step": m.e2e_traffic["h2d_bytes"], "d2h_bytes_per_step": m.e2e_traffic["d2h_bytes"]},
        "pipeline_roofline": {"bytes_per_bp": round(bpb, 1), "probes_per_bp": round(q, 3),
                              "achieved_GBps": round(bpb * m.total_bp / (step_ms * 1e-3) / 1e9, 1),
                              "frac_of_hbm_peak": round(bpb * m.total_bp / (step_ms * 1e-3) / 1e9 / peak, 4)},
        "stages_ms": {k: round(v, 4) for k, v in m.stages.items() if not k.startswith("upload:")},
        "upload_stages_ms": {k: round(v, 4) for k, v in m.stages.items() if k.startswith("upload:")},
        "sizes": dict(m.stats, alignment_columns=int(m.traffic["alignment_columns"])),
        "gpu_launches": int(m.launches),
    }


def roofline_of(m, cfg, peak, peak_kind):
    st = {"cols": m.traffic["alignment_columns"], "L": m.total_bp, "records": m.stats.get("records", 0),
          "groups": m.stats.get("groups", 0), "N": m.total_bp}
    kern = {k: v for k, v in m.stages.items() if k in STAGE_MODELS and v > 0}
    if not kern:
        return None
    rows = []
    for k, ms in kern.items():
        nl = max(1, int(round(m.stage_launches.get(k, 1))))
        n_jobs = max(1, cfg["contigs"])
        per_launch_ms = ms / n_jobs  # stage time summed over the contigs of the step; one stage instance per contig
        ach = STAGE_MODELS[k][1](st) / n_jobs / (per_launch_ms * 1e-3) / 1e9
        rows.append({"stage": k, "kernel": STAGE_MODELS[k][0], "achieved": round(ach, 1), "frac": round(ach / peak, 4),
                     "launch_ms": round(per_launch_ms, 4), "kernels_in_stage": nl,
                     "share_of_step": round(ms / m.stages["total"], 4),
                     "algorithmic_bytes_per_launch": int(STAGE_MODELS[k][1](st) / n_jobs), "traffic": NCU_TRAFFIC.get(k)})
    rows.sort(key=lambda r: -r["launch_ms"])
    dom = rows[0]
    return {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
            "frac": dom["frac"], "traffic": dom["traffic"], "peak_source": peak_kind, "launch_ms": dom["launch_ms"],
            "share_of_step": dom["share_of_step"], "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
            "stage": dom["stage"], "note": "the slowest stage of the step that has a byte model (DESIGN.md 4); the others follow",
            "other_stages": rows[1:]}


def cfg_for(n, full):
    cfg = dict(CONFIGS[n])
    if not full:
        cfg.update(cfg.get("bounded", {}))
    cfg.pop("bounded", None)
    return cfg


def bind_to_gpu_cpus(local, world):
    """One rank per GPU on a multi-socket box: keep this rank's threads (and, by first touch, its page-locked record
    buffers) on the CPUs NVML reports as local to its GPU, so that K0's reads of host memory and the record parse do not
    cross the socket interconnect.  A no-op when NVML reports no proper subset of the CPUs this process may use."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        mine = os.sched_getaffinity(0)
        cpus &= mine
        if cpus and cpus != mine and len(cpus) >= max(2, len(mine) // max(1, world)):
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "of": len(mine)}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:80]}
    return None


def run_ours(args):
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner ...) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    if line is not None:
        print(json.dumps(line), flush=True)


def _run_ours(args):
    import torch
    import torch.distributed as dist
    import nextpolish2_b200 as np2
    from nextpolish2_b200.api import set_host_threads
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    numa = bind_to_gpu_cpus(local, world) if world > 1 and not args.no_bind else None
    cores = max(1, (os.cpu_count() or 8) // max(world, 1))
    threads = min(cores, 16)
    if args.e2e_inflight <= 0:
        args.e2e_inflight = max(1, min(3, cores // 2))
    cfg = cfg_for(args.config, args.full or args.config == 1)
    if args.length:
        cfg["length"] = args.length
    if args.contigs:
        cfg["contigs"] = args.contigs
    if args.het is not None:
        cfg["het"] = args.het
    ctx = np2.Context(local)
    # the box's cores are shared by the ranks and by the contigs each rank keeps in flight
    set_host_threads(args.host_threads or max(2, min(16, cores // max(1, args.e2e_inflight))))
    opts_kw = {}
    opts = np2.Opts(**opts_kw)  # reference defaults; every contig is above -L 1000000

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    contigs, tabs = make_workload(cfg, SEED0 + args.config + 1000 * rank, threads)
    tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in cfg["ks"]]
    sampler = ClockSampler(local)
    m = measure(np2, torch, ctx, local, contigs, tables, cfg, args, opts, args.steps, args.warmup, barrier, args.e2e_inflight)
    clocks = sampler.stop()  # sampled over both timed regions (device-resident steps and end-to-end steps)

    t_dev = torch.tensor([m.dev_time, m.e2e_time, m.wall, m.e2e_serial_time], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    m.dev_time, m.e2e_time, m.wall, m.e2e_serial_time = [float(x) for x in t_dev.tolist()]

    ver = None
    if args.no_verify and not cfg["het"]:
        # size-independent property instead of the oracle run: polishing a haploid contig's reads gives back the haplotype
        # the reads were drawn from, byte for byte (tests/test_oracle_polish.py holds the oracle to the same property)
        ver = {"oracle": "skipped (--no-verify)", "contigs_checked": len(contigs),
               "identical_to_truth_haplotype": bool(all(bytes(r.split(b"\n", 2)[1]) == bytes(c["hap1"])
                                                        for r, c in zip(m.records, contigs)))}
    if not args.no_verify:
        ver = verify(contigs, tabs, cfg, m.records, m.dropped, opts_kw, cores)
        if world > 1:  # every rank checked its own contigs
            allv = [None] * world
            dist.all_gather_object(allv, ver)
            ver = dict(allv[0])
            ver["identical"] = all(v["identical"] for v in allv)
            ver["contigs_checked"] = sum(v["contigs_checked"] for v in allv)
            ver["per_rank_sha256_gpu"] = [v["fasta_sha256_gpu"][:16] for v in allv]

    line = None
    peak, peak_kind = peaks()
    if rank == 0:
        s = summarise(m, cfg, args.steps, world, peak)
        mbp_total = world * m.total_bp * args.steps / 1e6
        line = {
            "metric": "polished Mbp/s", "value": s["value"], "unit": "Mbp/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": s["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/u64 integer",
            "data": "synthetic",
            "config": {"workload": config_text(cfg), "baseline_config": cfg["label"], "contig_bp": cfg["length"],
                       "contigs_per_gpu": cfg["contigs"], "depth": cfg["depth"],
                       "tables": "+".join("k%d" % k for k in cfg["ks"]) + " synthesised from the truth haplotype(s)",
                       "partition": "one workload replica per GPU, no collective on the data path",
                       "l2": "inputs (%.0f MB of BAM records per GPU) are larger than the 126 MB L2" % (
                           sum(len(c["bam"]) for c in contigs) / 1e6)},
            "e2e": dict(s["e2e"], contigs_in_flight=args.e2e_inflight * (2 if args.e2e_prefetch and args.e2e_inflight > 1 else 1),
                        worker_threads=args.e2e_inflight, prefetch=bool(args.e2e_prefetch),
                        one_at_a_time=round(mbp_total / m.e2e_serial_time, 3),
                        one_at_a_time_ms={k: round(v / max(1, m.parts1["n"]), 3) for k, v in m.parts1.items()
                                          if k in ("create_parse", "upload", "run", "result")},
                        one_at_a_time_upload_stages_ms={k: round(v / max(1, m.parts1["n"]), 3)
                                                        for k, v in m.parts1.get("stages", {}).items() if k.startswith("upload:")},
                        in_flight_ms_per_contig_per_thread={k: round(v / max(1, m.partsN["n"]), 3)
                                                            for k, v in m.partsN.items() if k in ("create_parse", "upload", "run", "result")},
                        in_flight_stage_ms_per_contig={k: round(v / max(1, m.partsN["n"]), 3)
                                                       for k, v in sorted(m.partsN.get("stages", {}).items(), key=lambda kv: -kv[1])[:14]},
                        repeated_passes=m.partsN.get("repeated_passes", 0) + m.parts1.get("repeated_passes", 0)),
            "gpu_launches": s["gpu_launches"],
            "cpu_binding": numa,
            "clocks": clocks,
            "roofline": roofline_of(m, cfg, peak, peak_kind),
            "pipeline_roofline": s["pipeline_roofline"],
            "stages_ms": s["stages_ms"],
            "sizes": s["sizes"],
            "verify": ver,
            "identical_to_oracle": None if ver is None else ver.get("identical"),
            "wall_ms_per_step": round(m.wall / args.steps * 1e3, 3),
        }
    del m
    if rank == 0 and world == 1 and not args.no_bgzf_bench:  # the ingest step in front of the path, on the same records
        try:
            c0 = contigs[0]
            line["bgzf_inflate"] = bgzf_bench(ctx, np2, c0["bam"], int(len(c0["contig"])), cores, contig=c0["contig"],
                                              tables=tables, opts=opts, truth=c0["hap1"] if not len(c0["hap2"]) else None)
        except Exception as e:  # noqa: BLE001
            line["bgzf_inflate"] = {"error": repr(e)[:300]}
    for t in tables:
        t.free()
    del contigs, tabs

    # ---- the other configs, bounded (N = 1): same measurement, fewer steps
    if world == 1 and not args.no_extras:
        extras = {}
        for n in (2, 3):
            if n == args.config:
                continue
            ecfg = cfg_for(n, False)
            try:
                ec, et = make_workload(ecfg, SEED0 + n, threads)
                etab = [np2.Table.from_arrays(ctx, k, *et[k]) for k in ecfg["ks"]]
                em = measure(np2, torch, ctx, local, ec, etab, ecfg, args, opts, max(2, args.steps // 4), 3, barrier,
                             args.e2e_inflight)
                es = summarise(em, ecfg, max(2, args.steps // 4), 1, peak)
                es.pop("stages_ms")
                es.pop("upload_stages_ms")
                es["stages_ms_top"] = dict(sorted([kv for kv in em.stages.items() if not kv[0].startswith("upload:")],
                                                  key=lambda kv: -kv[1])[:10])
                es["stages_ms_top"] = {k: round(v, 3) for k, v in es["stages_ms_top"].items()}
                es["verify"] = None if args.no_verify else verify(ec, et, ecfg, em.records, em.dropped, opts_kw, cores)
                extras[str(n)] = es
                for t in etab:
                    t.free()
                del em, ec, et
            except Exception as e:  # noqa: BLE001
                extras[str(n)] = {"error": repr(e)[:300]}
        line["configs"] = extras

    if not args.no_strong:
        ss = strong_scaling(args, np2, torch, dist, ctx, local, rank, world, cores, barrier)
        if rank == 0:
            line["strong_scaling"] = ss

    if rank == 0:
        if not args.no_yak_bench:
            line["yak_lookup"] = yak_bench(ctx, np2, torch, peak)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(args, steps=args.cpu_steps)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


def strong_contig_lengths(total_bp, n=24):
    """configs[4]: 24 contigs of 50-250 Mbp summing to 3 Gbp, scaled to total_bp."""
    raw = np.linspace(250.0, 50.0, n)
    return [int(x / raw.sum() * total_bp) // 1000 * 1000 for x in raw]


def strong_scaling(args, np2, torch, dist, ctx, local, rank, world, cores, barrier):
    """configs[4] scaled: a FIXED set of 24 diploid contigs, LPT-partitioned over the ranks (contigs are independent:
    no collective on the data path); the tables are staged by rank 0 and broadcast over NVLink; every rank polishes its
    share end to end (host buffers -> FASTA records, several contigs in flight); rank 0 gathers the records in input
    order.  The digest must be the same at every N; rank 0 checks its own contigs against the oracle."""
    from nextpolish2_b200 import synth
    from nextpolish2_b200.shard import lpt_partition, broadcast_tables
    lens = strong_contig_lengths(int(args.strong_mbp * 1e6))
    seed, ks, threads = SEED0 + 4, (21, 31), min(cores, 16)
    opts_kw = {"min_ctg_len": 100000}  # the contigs are scaled down ~60x from configs[4]; so is -L
    opts = np2.Opts(**opts_kw)
    t0 = time.time()
    tabs, tables = None, []
    if rank == 0:  # haplotypes of every contig (no reads) -> tables
        haps = []
        for i, ln in enumerate(lens):
            A = synth.genome(seed + 10 * i, ln)
            c = synth.make_contig(seed + 10 * i + 1, A, depth=0.0, asm_err=2e-5, het=0.01, read_err=0.002, threads=1)
            haps += [c["hap1"], c["hap2"]]
        tabs = {k: synth.make_table_mt(seed + 2, k, haps, threads=threads) for k in ks}
        tables = [np2.Table.from_arrays(ctx, k, *tabs[k]) for k in ks]
        del haps
    t_tab = time.time() - t0
    torch.cuda.synchronize()
    t0 = time.time()
    tables = broadcast_tables(ctx, tables, rank, world)
    torch.cuda.synchronize()
    t_bcast = time.time() - t0
    mine = lpt_partition([float(x) for x in lens], world)[rank]
    t0 = time.time()
    contigs = []
    for i in mine:
        A = synth.genome(seed + 10 * i, lens[i])
        c = synth.make_contig(seed + 10 * i + 1, A, depth=30.0, asm_err=2e-5, het=0.01, read_err=0.002, threads=threads)
        contigs.append({"idx": i, "name": "ctg%03d" % i, "contig": A, "hap1": c["hap1"], "hap2": c["hap2"], "bam": c["bam"]})
    pinned = [(torch.from_numpy(c["contig"].copy()).pin_memory(), torch.from_numpy(c["bam"]).pin_memory()) for c in contigs]
    t_synth = time.time() - t0
    # small contigs are latency bound (a run is ~9 round trips and ~250 short kernels whatever the size): more workers
    # than the 3 that saturate the link with 10 Mbp contigs keep the GPU busier, as far as the rank's cores allow
    small = contigs and sum(len(c["contig"]) for c in contigs) / len(contigs) < 5e6
    inflight = max(1, min(max(args.e2e_inflight, min(6, cores // 2)) if small else args.e2e_inflight, len(contigs)))
    depth = 2 if (args.e2e_prefetch and len(contigs) > inflight) else 1  # see measure(): next contig parsed + uploading
    ctxs = [ctx] + [np2.Context(local) for _ in range(inflight * depth - 1)]

    def begin(cx, ci):
        p = pinned[ci]
        return np2.Job(cx, p[0].numpy(), p[1].numpy(), tables, opts), ci

    parts = {"upload_wait": 0.0, "run": 0.0, "result": 0.0, "n": 0, "repeated_passes": 0, "host_syncs": 0}

    def finish(started):
        j, ci = started
        t0 = time.perf_counter()
        j.upload()
        t1 = time.perf_counter()
        j.run(-1)
        t2 = time.perf_counter()
        first, last, base = j.bases(copy=False)
        rec = fasta_record(contigs[ci]["name"], first, last, base)  # header + a copy of the bases, in host memory
        nd = len(j.dropped())
        st = j.stats()
        j.destroy()
        t3 = time.perf_counter()
        for k, v in (("upload_wait", t1 - t0), ("run", t2 - t1), ("result", t3 - t2)):
            parts[k] += v * 1e3
        parts["n"] += 1
        parts["repeated_passes"] += st.get("repeated_passes", 0)
        parts["host_syncs"] += st.get("host_syncs", 0)
        return rec, nd

    def polish(cx, ci):
        return finish(begin(cx, ci))

    def one_pass():
        out, nxt, lock, errs = {}, [0], threading.Lock(), []

        def take():
            with lock:
                i = nxt[0]
                nxt[0] += 1
            return i if i < len(contigs) else None

        def work(w):
            try:
                mine, turn = ctxs[w * depth:(w + 1) * depth], 0
                i = take()
                cur = begin(mine[0], i) if i is not None else None
                while cur is not None:
                    nx = None
                    if depth > 1:
                        i = take()
                        if i is not None:
                            turn ^= 1
                            nx = begin(mine[turn], i)
                    out[contigs[cur[1]]["idx"]] = finish(cur)
                    if depth == 1:
                        i = take()
                        nx = begin(mine[0], i) if i is not None else None
                    cur = nx
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=work, args=(w,)) for w in range(inflight)]
        [t.start() for t in th]
        [t.join() for t in th]
        if errs:
            raise errs[0]
        return out
    from nextpolish2_b200.api import set_stage_timing
    set_stage_timing(False)
    if contigs:
        for cx in ctxs:
            polish(cx, 0)  # warm pools
        one_pass()  # untimed: every context has then seen contigs of several sizes (pools, speculative capacities)
    for k in parts:
        parts[k] = 0
    passes = max(1, args.strong_passes)
    barrier()
    t1 = time.perf_counter()
    for _ in range(passes):
        local_out = one_pass()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    barrier()
    for cx in ctxs[1:]:
        cx.close()
    set_stage_timing(True)
    t_dev = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dt_max = float(t_dev.item())
    digests = {i: (hashlib.sha256(r).hexdigest(), len(r), nd) for i, (r, nd) in local_out.items()}
    gathered = [digests]
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(digests, gathered, dst=0)
    res = None
    if rank == 0:
        merged = {}
        for d in gathered:
            merged.update(d)
        assert sorted(merged) == list(range(len(lens))), "a contig is missing from the gathered FASTA"
        h = hashlib.sha256()
        for i in range(len(lens)):  # input order (= the reference with -t 1)
            h.update(merged[i][0].encode())
        orec, odrop, osec = oracle_fasta(contigs, tabs, ks, opts_kw, cores) if not args.no_verify else ([], [], [])
        ok = all(hashlib.sha256(o).hexdigest() == merged[c["idx"]][0] for o, c in zip(orec, contigs))
        total = float(sum(lens))
        res = {"workload": "configs[4] scaled: 24 diploid contigs (1%% het, 30x), lengths %.2f-%.2f Mbp (configs[4]'s 250-50 Mbp "
                           "distribution), %.0f Mbp in total, -L 100000, k21+k31" % (max(lens) / 1e6, min(lens) / 1e6, total / 1e6),
               "n_gpus": world, "scaling": "strong", "partition": "LPT by contig length over the ranks, no data-path collective",
               "value": round(total * passes / 1e6 / dt_max, 3), "unit": "Mbp/s (end to end, host buffers in, FASTA records out)",
               "seconds_per_pass": round(dt_max / passes, 4), "passes": passes, "contigs_in_flight_per_gpu": inflight,
               "contigs_per_rank": [len(p) for p in lpt_partition([float(x) for x in lens], world)],
               "rank0_ms_per_contig_per_thread": {k: round(v / max(1, parts["n"]), 3) for k, v in parts.items() if k != "n"},
               "fasta_sha256_in_input_order": h.hexdigest(), "fasta_bytes": int(sum(v[1] for v in merged.values())),
               "reads_dropped_by_phasing": int(sum(v[2] for v in merged.values())),
               "oracle_checked_contigs": [c["idx"] for c in contigs] if not args.no_verify else [],
               "identical_to_oracle": (bool(ok) if not args.no_verify else None),
               "table_stage_s": round(t_tab, 2), "table_broadcast_s": round(t_bcast, 3),
               "table_bytes": int(sum(t.device_bytes for t in tables)),
               "table_broadcast_GBps": (round(sum(t.device_bytes for t in tables) / t_bcast / 1e9, 1) if world > 1 else None),
               "synth_s": round(t_synth, 1)}
    if rank != 0:  # adopted replicas
        for t in tables:
            t.free()
    else:
        for t in tables:
            t.free()
    return res


def bgzf_bench(ctx, np2, bam_records, ref_len, cores, contig=None, tables=None, opts=None, truth=None):
    """The ingest step in front of the hot path: the headline contig's records written as a BAM
    (BGZF level 1 and 6, the synthetic-input tool's writer), every member inflated on the device by np2_bgzf_inflate
    (a group of lanes per member) from a page-locked copy of the file into a page-locked record buffer.  Reported: the
    kernel's time from CUDA events, the whole call with the compressed span going up and the records coming back, the
    same members through zlib on the host threads (ThreadPool; zlib releases the GIL), and the byte-for-byte comparison
    with zlib's output (oracle/bgzf.py)."""
    import tempfile
    import zlib
    from concurrent.futures import ThreadPoolExecutor
    from nextpolish2_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bgzf as OB  # the checker: member walk + zlib
    res = {}
    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    for level in (1, 6):
        path = os.path.join(d, "np2_bench_%d.l%d.bam" % (os.getpid(), level))
        synth.write_bam(path, ["ctg000"], [ref_len], [bam_records], level=level)
        buf = np.fromfile(path, np.uint8)
        for f in (path, path + ".bai"):
            if os.path.exists(f):
                os.remove(f)
        po, pl, iz = np2.bgzf_members(buf)
        total = int(iz.astype(np.uint64).sum())
        raw = bytes(buf)

        def host(i):
            return zlib.decompress(raw[int(po[i]):int(po[i]) + int(pl[i])], -15)
        with ThreadPoolExecutor(cores) as ex:
            t0 = time.perf_counter()
            parts = list(ex.map(host, range(len(po)), chunksize=64))
            t_host = time.perf_counter() - t0
        want = b"".join(parts)
        del parts
        pin_in, pin_out = np2.PinnedBuffer(buf), np2.PinnedBuffer(np.zeros(total, np.uint8))
        ks, ws = [], []
        for rep in range(6):
            t0 = time.perf_counter()
            got, kms = np2.bgzf_inflate(ctx, pin_in, po, pl, iz, out=pin_out)
            ws.append((time.perf_counter() - t0) * 1e3)
            ks.append(kms)
        same = bytes(got) == want
        kms, wms = float(np.median(ks[1:])), float(np.median(ws[1:]))
        res["level%d" % level] = {
            "members": int(len(po)), "compressed_MB": round(len(buf) / 1e6, 1), "records_MB": round(total / 1e6, 1),
            "kernel_ms": round(kms, 3), "kernel_GBps_out": round(total / kms / 1e6, 1), "kernel_GBps_in": round(len(buf) / kms / 1e6, 1),
            "call_ms_pinned_in_pinned_out": round(wms, 2), "h2d_bytes": int(len(buf)), "d2h_bytes": total,
            "host_zlib_ms": round(t_host * 1e3, 1), "host_threads": cores, "identical_to_zlib": bool(same)}
        if level == 1 and tables is not None:
            # file bytes -> polished bases through np2_job_create_bgzf (records inflated, walked and gathered on the
            # device; only the record heads come down), one contig at a time on one context: what the command line
            # runs per lane.  The result must be the contig the record path produces (= the truth haplotype here).
            skip = want.find(bytes(bam_records[:4096]))
            ref_job = np2.Job(ctx, contig, bam_records, tables, opts).upload().run()
            ref_base = bytes(ref_job.bases()[2])
            ref_job.destroy()
            ts, same_job = [], None
            for rep in range(4):
                t0 = time.perf_counter()
                job = np2.Job.from_bgzf(ctx, contig, pin_in, po, pl, iz, skip, len(bam_records), tables, opts).upload().run()
                first, last, base = job.bases(copy=False)
                n_base = len(base)
                ts.append((time.perf_counter() - t0) * 1e3)
                if rep == 0:
                    same_job = bool(bytes(base) == ref_base)
                    same_truth = None if truth is None else bool(ref_base == bytes(truth))
                path = job.ingest_path
                job.destroy()
            res["job_from_members"] = {
                "entry": "np2_job_create_bgzf -> np2_job_upload -> np2_job_run -> np2_job_get_consensus",
                "ms_per_contig_one_at_a_time": round(float(np.median(ts[1:])), 2),
                "Mbp_per_s_one_at_a_time": round(ref_len / 1e3 / float(np.median(ts[1:])), 1),
                "h2d_bytes": int(len(buf)) + ref_len, "ingest_path": int(path),
                "identical_to_job_from_records": same_job, "identical_to_truth_haplotype": same_truth,
                "note": "three lanes of the command line overlap these steps: 11 ms per contig (profiles/r02az_cli_e2e.txt)"}
        pin_in.free()
        pin_out.free()
        del want, raw, buf
    res["kernel"] = "k_bgzf_inflate<16, 4> (np2_inflate.cu): 16 lanes per BGZF member, Huffman tables in shared memory"
    res["note"] = ("not part of `value` / `e2e` (those start from uncompressed records, as the reference's worker closure does); "
                   "the command line's file -> FASTA numbers (np2_job_create_bgzf: the records stay on the device) are in profiles/r02az_cli_e2e.txt")
    return res


def yak_bench(ctx, np2, torch, peak):
    """K5 at scale: 2^26 keys (~0.9 GB table >> L2), 2^26 device-resident queries, half present / half absent.
    One probe needs one 32-byte bucket.  What DRAM really delivers per L2 miss depends on the device's L2 fetch
    granularity (cudaLimitMaxL2FetchGranularity): the lookup is measured at 32 / 64 / 128 bytes, next to a plain random
    gather of whole 32 / 64 / 128-byte blocks (the peak of each access size), and the best setting is reported."""
    from nextpolish2_b200.api import bench_gather, l2_fetch_granularity
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
    probes = q.numel()
    default_gran = l2_fetch_granularity(ctx, 0)
    sweep = {}
    for gran in (32, 64, 128):
        l2_fetch_granularity(ctx, gran)
        tab.lookup_device(q.data_ptr(), probes, out.data_ptr(), 5, repeat=3)
        ms = tab.lookup_device(q.data_ptr(), probes, out.data_ptr(), 5, repeat=10)
        g32 = bench_gather(ctx, tab.device_bytes, probes, 32, repeat=10)
        sweep[str(gran)] = {"lookup_ms": round(ms, 4), "gprobes_per_s": round(probes / ms / 1e6, 3),
                            "gather32_GBps": round(probes * 32 / g32 / 1e6, 1)}
    l2_fetch_granularity(ctx, default_gran)
    blocks = {}
    for bb in (32, 64, 128):  # whole aligned blocks of bb bytes at the default granularity: the DRAM-side peak per size
        gm = bench_gather(ctx, tab.device_bytes, probes, bb, repeat=10)
        blocks[str(bb)] = {"Gblocks_per_s": round(probes / gm / 1e6, 3), "GBps": round(probes * bb / gm / 1e6, 1)}
    best = min(sweep, key=lambda k: sweep[k]["lookup_ms"])
    ms = sweep[best]["lookup_ms"]
    gather_gbs = max(v["gather32_GBps"] for v in sweep.values())
    # size-independent check: present keys answer their count (>= 5 filter), absent answer 0
    exp = torch.where(cnt[perm] >= 5, cnt[perm], torch.zeros_like(cnt[perm]))
    ok = bool((tab.lookup(keys[perm].cpu().numpy().view(np.uint64), 5) == exp.cpu().numpy().view(np.uint16)).all())
    res = {"probes": probes, "table_keys": nk, "table_bytes": tab.device_bytes, "ms": ms,
           "gprobes_per_s": round(probes / ms / 1e6, 3),
           "sector_GBps": round(probes * 32 / ms / 1e6, 1), "algorithmic_GBps": round(probes * 42 / ms / 1e6, 1),
           "frac_of_stream_peak": round(probes * 42 / ms / 1e6 / peak, 4),
           "random_gather_peak_GBps": gather_gbs, "frac_of_random_gather_peak": round(probes * 32 / ms / 1e6 / gather_gbs, 4),
           "l2_fetch_granularity": {"default": default_gran, "best": int(best), "sweep": sweep},
           "random_block_gather": blocks,
           "bucket_load": "one 256-bit ld.global.nc.L2::evict_first.L2::64B per bucket (LDG.E.EFL2.LTC64B.256)",
           "dram_bytes_per_probe_ncu": {"value": 83.9, "was": 153.6, "source": "profiles/r02fin2_yak_probe_full.txt (ncu --set full of "
                                        "this arm); random 32-byte loads run at the same ~43 G/s whether a miss fills 64 or 128 "
                                        "bytes (profiles/r02_l2_fetch.txt): the bound is DRAM activations, not bytes"},
           "correct": ok}
    tab.free()
    return res


def cpu_sample(args, steps=1, warmup=0, faithful=False):
    """The reference's CPU algorithm (oracle port) on a bounded sample of configs[1]: `cores` contigs of 1 Mbp each, one
    worker thread per contig (the reference's unit of parallelism, main.rs:1698-1853).  faithful: every k-mer retrieval
    pass scans ALL keys of the table against the query set (kmer.rs:132-170: the reference streams the whole .yak file
    per pass, per contig, per thread) instead of probing an in-memory hash table."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from nextpolish2_b200 import synth
    cfg = CONFIGS[1]
    cores = min(os.cpu_count() or 1, 32)
    clen = args.cpu_contig
    G = synth.genome(SEED0 + 2, clen * cores)
    data = []
    for i in range(cores):
        A = G[i * clen:(i + 1) * clen].copy()
        data.append((A, synth.make_contig(SEED0 + 100 + i, A, depth=cfg["depth"], asm_err=2e-5, het=0.0, read_err=0.002, threads=4)))
    haps = [d[1]["hap1"] for d in data]
    tabs = [O.Table.from_arrays(k, *synth.make_table_mt(SEED0 + 3, k, haps, threads=min(cores, 16))) for k in cfg["ks"]]
    for t in tabs:
        t.set_stream_scan(faithful)
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
        assert all(bytes(res[i].consensus()[1]) == bytes(haps[i]) for i in range(cores))
        if it >= warmup:
            vals.append(cores * clen / 1e6 / dt)
    how = ("every retrieval pass scans all table keys against the query set, as the reference streams the whole .yak file "
           "per pass (kmer.rs:132-170); keys held in memory, so still kinder than the reference's file reads"
           if faithful else "tables probed in memory (kinder to the CPU than the reference's per-pass .yak file streaming)")
    return {"value": round(float(np.mean(vals)), 4), "unit": "Mbp/s", "cores": cores, "kind": "port",
            "sample": "%d contigs x %d bp of configs[1]'s synthetic workload (30x), one oracle thread per contig, %s; "
                      "restatement of the reference algorithm, not the Rust binary (no cargo in this image)" % (cores, clen, how),
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
            "config": {"workload": CONFIGS[1]["text"], "baseline_config": "configs[1]", "sample": b["sample"]},
            "cpu_baseline": {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": b["value"], "unit": "Mbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not args.no_ref_faithful:
        f = cpu_sample(args, steps=max(1, min(args.steps, 3)), warmup=0, faithful=True)
        line["ref_faithful"] = {k: f[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3], help="BASELINE.json configs[N] as the headline workload")
    ap.add_argument("--full", action="store_true", help="configs 2/3 at their stated size (10 x 10 Mbp; 250 Mbp)")
    ap.add_argument("--length", type=int, default=0, help="override the contig length")
    ap.add_argument("--contigs", type=int, default=0, help="override the number of contigs per GPU")
    ap.add_argument("--het", type=float, default=None, help="override the heterozygosity")
    ap.add_argument("--cpu-contig", type=int, default=1_000_000)
    ap.add_argument("--cpu-steps", type=int, default=10, help="passes over the CPU sample for cpu_baseline (~1 s each)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-yak-bench", action="store_true")
    ap.add_argument("--no-bgzf-bench", action="store_true", help="skip the device BGZF inflate arm")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle FASTA comparison")
    ap.add_argument("--no-extras", action="store_true", help="skip the bounded configs[2] / configs[3] arms")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling arm (configs[4] scaled)")
    ap.add_argument("--no-ref-faithful", action="store_true")
    ap.add_argument("--strong-mbp", type=float, default=36.0, help="total size of the 24 contigs of the strong-scaling arm")
    ap.add_argument("--strong-passes", type=int, default=3)
    ap.add_argument("--pageable", action="store_true", help="record buffer in pageable memory (host compaction path)")
    ap.add_argument("--host-threads", type=int, default=0, help="host threads per library call (0 = cores / ranks / ~in flight)")
    ap.add_argument("--e2e-inflight", type=int, default=0,
                    help="worker threads per GPU in the end-to-end arm (0 = min(3, host cores per rank / 2): 3 on a 16-core box "
                         "with one GPU, 2 with 8 ranks on 32 cores, where a third worker only oversubscribes the cores)")
    ap.add_argument("--no-bind", action="store_true", help="N > 1: do not bind the rank to the CPUs local to its GPU")
    ap.add_argument("--e2e-stage-timers", action="store_true", help="keep the per-stage event timers on with contigs in flight")
    ap.add_argument("--e2e-prefetch", type=int, default=1, help="1: every worker parses and starts the upload of its next contig before it runs the current one")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
