#!/usr/bin/env python
"""End-to-end probe of the BAM route:  python tools/e2e_probe.py [reads] [streams,...]
compressed BAM bytes in pinned host memory -> wgbs_dbam_open -> wgbs_pileup_dbam -> pat2beta -> collapse -> pat text + .beta in pinned
host memory, with S batches in flight (one Context / stream / host thread each).  One JSON line per S: ms per step (CUDA events
spanning all streams), reads/s, per-kernel breakdown of one profiled step."""
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (workload generators)


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    S_list = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,3,4").split(",")]
    steps, warmup = 12, 3
    from wgbs_tools_b200 import bamio
    sam = bench.make_batch(n_reads, 1000)
    n_rec = sam.count(10)
    t0 = time.time()
    bam = bamio.sam_to_bam(sam, [(bench.CHR, bench.CHR_LEN)], procs=max(1, min(bench.host_threads(), 32)))
    bench.log(f"[probe] BAM {len(bam) / 1e6:.1f} MB ({time.time() - t0:.1f}s)")
    import torch
    from wgbs_tools_b200._lib import PileupOpts, ViewOpts, check, lib
    from wgbs_tools_b200.api import Context
    g = bench.genome(); n_cpg = g.n_cpg
    torch.cuda.set_device(0)
    main_stream = torch.cuda.Stream(); torch.cuda.set_stream(main_stream)
    ctx0 = Context(0, stream=main_stream.cuda_stream)
    ix = ctx0.load_index(g.loci, 1)
    h_bam = torch.frombuffer(bytearray(bam), dtype=torch.uint8).pin_memory()
    ref = {}
    for S in S_list:
        streams = [main_stream] if S == 1 else [torch.cuda.Stream() for _ in range(S)]
        ctxs = [ctx0] if S == 1 else [Context(0, stream=st.cuda_stream) for st in streams]
        h_text = [torch.empty(32 << 20, dtype=torch.uint8).pin_memory() for _ in range(S)]
        h_beta = [torch.empty((n_cpg, 2), dtype=torch.uint8).pin_memory() for _ in range(S)]
        mc = [torch.zeros((n_cpg, 2), dtype=torch.int32, device="cuda") for _ in range(S)]
        out_n = [{} for _ in range(S)]; errs = []
        torch.cuda.synchronize()

        laps = {}

        def lap(name, t0):
            laps[name] = laps.get(name, 0.0) + (time.perf_counter() - t0)
            return time.perf_counter()

        def step(w):
            ctx = ctxs[w]
            B = C.c_void_p()
            t = time.perf_counter()
            check(lib.wgbs_dbam_open(ctx.h, h_bam.data_ptr(), h_bam.numel(), C.byref(B)))
            t = lap("dbam_open", t)
            vo = ViewOpts(); vo.refid = 0
            o = PileupOpts(1, 0, -1, 0, 0, 0.67, b"C")
            h = C.c_void_p(); st = (C.c_uint64 * 8)()
            check(lib.wgbs_pileup_dbam(ctx.h, ix.h, B, C.byref(vo), C.addressof(o), C.byref(h), C.addressof(st), None))
            t = lap("pileup_dbam", t)
            lib.wgbs_dbam_close(ctx.h, B)
            check(lib.wgbs_pat2beta(ctx.h, h, 1, n_cpg + 1, mc[w].data_ptr(), 1))
            t = lap("close+pat2beta", t)
            check(lib.wgbs_collapse(ctx.h, h))
            t = lap("collapse", t)
            n = C.c_size_t()
            check(lib.wgbs_pats_format(ctx.h, h, bench.CHR.encode(), h_text[w].data_ptr(), h_text[w].numel(), C.byref(n)))
            t = lap("format", t)
            check(lib.wgbs_trim(ctx.h, mc[w].data_ptr(), n_cpg, 8, h_beta[w].data_ptr()))
            lib.wgbs_pats_free(ctx.h, h)
            t = lap("trim+free", t)
            out_n[w].update(n=n.value, lines=int(st[0]))

        def work(w, k):
            try:
                torch.cuda.set_device(0)
                for _ in range(k):
                    step(w)
            except Exception as e:
                errs.append(repr(e))

        def run(total):
            if S == 1:
                return work(0, total)
            th = [threading.Thread(target=work, args=(w, total // S + (1 if w < total % S else 0))) for w in range(S)]
            [t.start() for t in th]; [t.join() for t in th]

        run(warmup * S)
        torch.cuda.synchronize()
        if errs:
            print(json.dumps({"S": S, "error": errs[0]})); continue
        outs = [(h_text[w][:out_n[w]["n"]].numpy().tobytes(), h_beta[w].numpy().tobytes()) for w in range(S)]
        ref.setdefault("out", outs[0])
        same = all(o == ref["out"] for o in outs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(main_stream)
        for st in streams:
            if st is not main_stream:
                st.wait_event(e0)
        laps.clear()
        run(steps * S)
        res_laps = {k: round(v * 1e3 / (steps * S), 3) for k, v in laps.items()}       # host wall per call with S steps in flight
        for st in streams:
            if st is not main_stream:
                ev = torch.cuda.Event(); ev.record(st); main_stream.wait_event(ev)
        e1.record(main_stream)
        torch.cuda.synchronize()
        wall = time.time() - t0
        ms = e0.elapsed_time(e1) / (steps * S)
        res = {"S": S, "ms_per_step": ms, "wall_ms_per_step": wall * 1e3 / (steps * S), "reads_per_sec": n_rec / (ms / 1e3), "same_outputs": bool(same), "lines": out_n[0]["lines"],
               "text_bytes": out_n[0]["n"], "bam_bytes": len(bam), "host_wall_ms_per_call_in_flight": res_laps}
        if S == 1:
            laps.clear()
            for _ in range(4):
                step(0)
            res["host_wall_ms_per_call"] = {k: round(v * 1e3 / 4, 3) for k, v in laps.items()}
            ctx0.prof(True)
            for _ in range(2):
                step(0)
            rep = ctx0.prof_report()
            ctx0.prof(False)
            res["kernels_ms"] = {k: round(v[1] / 2, 4) for k, v in sorted(rep.items(), key=lambda kv: -kv[1][1])[:24]}
            res["kernel_sum_ms"] = round(sum(v[1] for v in rep.values()) / 2, 4)
        print(json.dumps(res), flush=True)
        if S > 1:
            for c in ctxs:
                c.close()


if __name__ == "__main__":
    main()
