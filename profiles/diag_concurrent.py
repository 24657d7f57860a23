"""How well do np2_job_run calls of several contexts share one GPU?  configs[1]'s contig resident on W contexts, W
threads each looping np2_job_run (no parse, no upload): ms per contig against the single-context step.
usage (GPU box): python profiles/diag_concurrent.py [workers ...]"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch
    import nextpolish2_b200 as np2
    from nextpolish2_b200.api import set_host_threads
    workers = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 4]
    cfg = bench.cfg_for(1, True)
    contigs, tabs = bench.make_workload(cfg, bench.SEED0 + 1, 16)
    ctx0 = np2.Context(0)
    tables = [np2.Table.from_arrays(ctx0, k, *tabs[k]) for k in cfg["ks"]]
    opts = np2.Opts()
    c = contigs[0]
    ctg = torch.from_numpy(c["contig"].copy()).pin_memory()
    bam = torch.from_numpy(c["bam"]).pin_memory()
    set_host_threads(4)
    out = {}
    for W in workers:
        ctxs = [ctx0] + [np2.Context(0) for _ in range(W - 1)]
        jobs = [np2.Job(cx, ctg.numpy(), bam.numpy(), tables, opts).upload() for cx in ctxs]
        for j in jobs:
            for _ in range(4):
                j.run(-1)
        torch.cuda.synchronize()
        steps = 20
        acc = [dict() for _ in range(W)]

        def work(w):
            for _ in range(steps):
                jobs[w].run(-1)
                for k, v in jobs[w].timings().items():
                    acc[w][k] = acc[w].get(k, 0.0) + v[0]
        th = [threading.Thread(target=work, args=(w,)) for w in range(W)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        [t.join() for t in th]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tot = {}
        for a in acc:
            for k, v in a.items():
                tot[k] = tot.get(k, 0.0) + v / (steps * W)
        top = dict(sorted(tot.items(), key=lambda kv: -kv[1])[:12])
        out[str(W)] = {"ms_per_contig": round(dt * 1e3 / (steps * W), 3), "ms_per_run_per_thread": round(dt * 1e3 / steps, 3),
                       "stage_ms_per_run": {k: round(v, 3) for k, v in top.items()}}
        for j in jobs:
            j.destroy()
        for cx in ctxs[1:]:
            cx.close()
    # mix: R threads looping np2_job_run on resident jobs, U threads looping np2_job_create + np2_job_upload (parse, K0)
    for R_, U_ in (((0, 1), (0, 2), (3, 1), (3, 2), (2, 2)) if not os.environ.get("DIAG_SKIP_MIX") else ()):
        ctxs = [np2.Context(0) for _ in range(R_ + U_)]
        jobs = [np2.Job(cx, ctg.numpy(), bam.numpy(), tables, opts).upload() for cx in ctxs[:R_]]
        for j in jobs:
            for _ in range(4):
                j.run(-1)
        for cx in ctxs[R_:]:
            for _ in range(3):
                np2.Job(cx, ctg.numpy(), bam.numpy(), tables, opts).upload().destroy()
        torch.cuda.synchronize()
        stop = [False]
        n_run, n_up, t_parse = [0] * R_, [0] * U_, [0.0] * U_

        def runner(w):
            while not stop[0]:
                jobs[w].run(-1)
                n_run[w] += 1

        def uploader(w):
            while not stop[0]:
                t0 = time.perf_counter()
                j = np2.Job(ctxs[R_ + w], ctg.numpy(), bam.numpy(), tables, opts)
                t_parse[w] += time.perf_counter() - t0
                j.upload()
                j.destroy()
                n_up[w] += 1
        th = [threading.Thread(target=runner, args=(w,)) for w in range(R_)] + [threading.Thread(target=uploader, args=(w,)) for w in range(U_)]
        t0 = time.perf_counter()
        [t.start() for t in th]
        time.sleep(0.5)
        stop[0] = True
        [t.join() for t in th]
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        out["mix_R%d_U%d" % (R_, U_)] = {"run_ms_per_contig": round(dt / max(1, sum(n_run)), 3) if R_ else None,
                                         "upload_ms_per_contig": round(dt / max(1, sum(n_up)), 3),
                                         "parse_ms_per_call": round(sum(t_parse) * 1e3 / max(1, sum(n_up)), 3)}
        for j in jobs:
            j.destroy()
        for cx in ctxs:
            cx.close()
    # what exactly does a saturated K0 slow down?  U uploader threads against (a) a device-to-device copy (HBM bound, no
    # host interaction), (b) a chain of tiny kernels without synchronisation (launch path), (c) tiny kernel + .item()
    # round trips (synchronisation path)
    a = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    b = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
    z = torch.zeros(64, device="cuda")

    def probe():
        r = {}
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(50):
                a.copy_(b)
            e1.record(st)
            st.synchronize()
            r["d2d_GBps"] = round(50 * 2 * a.numel() / e0.elapsed_time(e1) / 1e6, 1)
            t0 = time.perf_counter()
            for _ in range(2000):
                z.add_(1.0)
            st.synchronize()
            r["tiny_kernel_us"] = round((time.perf_counter() - t0) * 1e6 / 2000, 2)
            t0 = time.perf_counter()
            for _ in range(300):
                z.add_(1.0)
                z[0].item()
            r["kernel_plus_readback_us"] = round((time.perf_counter() - t0) * 1e6 / 300, 2)
        return r
    out["probe_idle"] = probe()
    for U_ in (1, 2, 3):
        ctxs = [np2.Context(0) for _ in range(U_)]
        for cx in ctxs:
            for _ in range(3):
                np2.Job(cx, ctg.numpy(), bam.numpy(), tables, opts).upload().destroy()
        stop = [False]

        def uploader(w):
            while not stop[0]:
                np2.Job(ctxs[w], ctg.numpy(), bam.numpy(), tables, opts).upload().destroy()
        th = [threading.Thread(target=uploader, args=(w,)) for w in range(U_)]
        [t.start() for t in th]
        time.sleep(0.05)
        out["probe_U%d" % U_] = probe()
        stop[0] = True
        [t.join() for t in th]
        for cx in ctxs:
            cx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
