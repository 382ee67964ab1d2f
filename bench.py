#!/usr/bin/env python
"""bench.py -- genotype calls/sec of the VCF -> tensor hot path on B200.

A "step" = one pass of the whole device-resident path over one synthetic VCF body resident in HBM:
record location (head walker / tokenizer) -> site extraction -> GT decode (kernels 1-3, output in the
byte-shuffled planar layout) -> kernel 4 (site templates, allele-plane LZ4, size scan, frame assembly):
decompressed text in, Blosc2 frames of every (donor, HDF5 chunk) out -- SURVEY.md 8(d)'s timing span.
`--parse-only` times kernels 1-3 alone.  N=1 workload:
BASELINE.json configs[1] (synthetic chr22-like, 1.1M biallelic variants x 2504 phased samples,
~11.1 GB of text).  N>1: every rank parses its own chromosome-sized shard of that shape (the
path shards by chromosome / BGZF range with no collective on the data path); a 6-int64
all_gather of per-shard index metadata is the only exchange.

Contract: python bench.py --gpus N --steps K --warmup W [--impl reference]; rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NOMINAL_HBM_GBS = 7700.0
FALLBACK_HBM_GBS = 6650.0


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(text_with_header: bytes, sample_names, n_variants_sample: int, threads: int, passes_per_thread: int = 1):
    """The reference's driving pattern (vcf_to_h5.py:150-152,191-192): one whole-text load_vcf pass per
    donor, `threads` donors in parallel.  Runs the C oracle (kind "port": the reference parser needs
    htslib and cannot be built here).  Returns genotype calls/s = donor passes * variants / wall."""
    import oracle
    oracle.lib()
    donors = [sample_names[(k * 7919) % len(sample_names)] for k in range(threads * passes_per_thread)]
    done = []

    def work(k):
        for j in range(passes_per_thread):
            r = oracle.parse_text(text_with_header, donors[k * passes_per_thread + j], "")
            done.append(r["n"])

    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
    [t.start() for t in th]
    [t.join() for t in th]
    dt = time.perf_counter() - t0
    calls = sum(done)
    return calls / dt, dt, len(donors)


def run_reference(args):
    """--impl reference: the reference path's CPU implementation (oracle port) on the host cores, on a
    bounded sample of the same workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from haplohyped_varawareml_b200 import capi
    cores = os.cpu_count() or 1
    nv = args.cpu_sample_variants
    spec = capi.synth_spec(nv, args.samples, seed=args.seed)
    text = capi.synth_header(spec) + capi.synth_host(spec)
    names = capi.synth_sample_names(spec)
    for _ in range(args.warmup):
        cpu_baseline(text, names, nv, cores)
    t_tot, calls_tot = 0.0, 0.0
    for _ in range(args.steps):
        v, dt, nd = cpu_baseline(text, names, nv, cores)
        t_tot += dt
        calls_tot += v * dt
    value = calls_tot / t_tot
    sample = (f"first {nv} variants x {args.samples} samples of the workload text ({len(text) / 1e6:.0f} MB); per step "
              f"{cores} donors parsed in parallel, one whole-text load_vcf pass per donor (reference driving pattern)")
    line = {"impl": "reference", "metric": "genotype calls/sec", "value": value, "unit": "calls/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": "calls/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "calls/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "oracle C port of parse_vcf.cpp:30-71 semantics; the reference binary needs htslib (absent)"}
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"synthetic chr22-like VCF body: {args.variants} biallelic SNPs x {args.samples} phased samples, "
                        f"FORMAT=GT, seed {args.seed} (BASELINE.json configs[1])",
            "variants": args.variants, "samples": args.samples,
            "l2": "inputs (GBs of text) are far larger than the 126 MB L2; no flush needed",
            "per_rank": "each rank parses its own shard of this shape (seed + rank)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--variants", type=int, default=1_100_000)
    ap.add_argument("--samples", type=int, default=2504)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--cpu-sample-variants", type=int, default=60000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-bgzf", action="store_true", help="skip the e2e variant that starts from BGZF bytes")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--slab-bytes", type=int, default=1 << 30, help="slab size of the streamed e2e path")
    ap.add_argument("--parse-only", action="store_true", help="step = kernels 1-3 only (no Blosc2 frames)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from haplohyped_varawareml_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # rank 0 prints ONE line on stdout: whatever libraries print there meanwhile (NCCL's version banner, when NCCL_DEBUG
    # is set in the environment) goes to stderr -- file descriptor 1 is pointed at stderr until the JSON line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    V, S = args.variants, args.samples
    spec = capi.synth_spec(V, S, seed=args.seed + rank, chrom="chr22")
    T = int(capi.lib().hb_synth_body_bytes(spec))
    text = torch.empty(T + 256, dtype=torch.uint8, device=dev)
    text[T:].zero_()
    capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, local, None))
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    meta = torch.zeros(6, dtype=torch.int64, device=dev)
    gathered = [torch.zeros(6, dtype=torch.int64, device=dev) for _ in range(world)]

    frames = [None]

    def step(p):
        p.rerun()
        if not args.parse_only:
            if frames[0] is None:
                frames[0] = p.compress(0)           # first call allocates (warm-up)
            else:
                frames[0].rerun(p)
        if world > 1:            # the path's only exchange: per-shard index metadata
            i = p.info
            meta[0], meta[1], meta[2] = int(i.n_records), int(i.n_lines), int(i.text_bytes)
            dist.all_gather(gathered, meta)

    # first parse allocates; it is warm-up step 1
    p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22", device=local, stream=stream)
    if not args.parse_only:
        frames[0] = p.compress(0)
        p.attach(frames[0])                         # site templates are made while the GT decoder runs
    for _ in range(args.warmup - 1):
        step(p)
    info = p.info
    Vk = int(info.n_records)

    barrier()
    launches0 = capi.kernel_launches()
    clocks = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tok, sit, dec = [], [], []
    k4 = {"site_templates": [], "donor_frames": []}
    ev0.record()
    for _ in range(args.steps):
        step(p)
        i = p.info
        tok.append(i.ms_tokenize); sit.append(i.ms_sites); dec.append(i.ms_decode)
        if frames[0] is not None:
            fi = frames[0].info
            k4["site_templates"].append(fi.ms_site); k4["donor_frames"].append(fi.ms_frames)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if clocks else None
    launches = capi.kernel_launches() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    calls_total = float(V) * S * world
    value = calls_total / (ms / 1e3 / args.steps)

    # ---- parity spot check of what was just measured (oracle as the checker, small sample)
    parity = None
    if rank == 0:
        import oracle
        nchk = 64
        head = capi.synth_header(spec) + capi.synth_host(spec, 0, nchk)
        ora = oracle.parse_text(head, "*", "chr22")
        start, stop, ref, alt = p.sites()
        ok = np.array_equal(start[:ora["n"]], ora["start"])
        for s in (0, S // 2, S - 1):
            g0, g1 = p.sample(s)
            ok = ok and np.array_equal(g0[:ora["n"]], ora["gt0"][s]) and np.array_equal(g1[:ora["n"]], ora["gt1"][s])
        parity = bool(ok)

    # ---- e2e: host (pinned) text -> C ABI -> genotype matrix back in host memory
    e2e = None
    try:
      if not args.no_e2e:
          # every rank must get its pinned buffers, or none runs the e2e leg (a rank that dropped out alone would leave
          # the others waiting at the barrier)
          alloc_err = None
          try:
              host = torch.empty(T, dtype=torch.uint8).pin_memory()
              host.copy_(text[:T])
              out0 = torch.empty((S, Vk), dtype=torch.int8).pin_memory()
              out1 = torch.empty((S, Vk), dtype=torch.int8).pin_memory()
              sites = [torch.empty(Vk, dtype=torch.int32).pin_memory(), torch.empty(Vk, dtype=torch.int32).pin_memory(),
                       torch.empty(Vk, dtype=torch.uint8).pin_memory(), torch.empty(Vk, dtype=torch.uint8).pin_memory()]
          except Exception as ex:
              alloc_err = "%s: %s" % (type(ex).__name__, ex)
          okf = torch.tensor([0 if alloc_err else 1], dtype=torch.int32, device=dev)
          if world > 1:
              dist.all_reduce(okf, op=dist.ReduceOp.MIN)
          if int(okf.item()) == 0:
              raise RuntimeError("pinned host buffers for the e2e leg could not be allocated on every rank (%s)" % (alloc_err or "another rank"))

          nrec = C.c_uint64()
          opts = capi.Parse._opts(S, "chr22", False, True, local, 0, None)

          def e2e_step():
              capi.check(capi.lib().hb_parse_stream_host(host.data_ptr(), T, C.byref(opts), args.slab_bytes, out0.data_ptr(),
                                                         out1.data_ptr(), Vk, sites[0].data_ptr(), sites[1].data_ptr(),
                                                         sites[2].data_ptr(), sites[3].data_ptr(), None, None, C.byref(nrec), None))
              assert nrec.value == Vk

          e2e_step()
          barrier()
          t0 = time.perf_counter()
          for _ in range(args.e2e_steps):
              e2e_step()
          barrier()
          dt = time.perf_counter() - t0
          tt = torch.tensor([dt], dtype=torch.float64, device=dev)
          if world > 1:
              dist.all_reduce(tt, op=dist.ReduceOp.MAX)
          dt = float(tt.item())
          e2e = {"value": calls_total / (dt / args.e2e_steps), "unit": "calls/s", "h2d_bytes_per_step": T * world,
                 "d2h_bytes_per_step": (2 * S * Vk + 10 * Vk) * world, "steps": args.e2e_steps,
                 "api": "hb_parse_stream_host: pinned host text -> genotype matrix [S][V'] x2 + site columns in pinned host "
                        "memory; slabs of %d MiB, H2D / kernels / D2H overlapped" % (args.slab_bytes >> 20)}
          # the streamed result is the same matrix the device-resident step produced
          if rank == 0:
              g0, g1 = p.sample(S // 3)
              e2e["matches_device_path"] = bool(np.array_equal(out0[S // 3].numpy(), g0) and np.array_equal(out1[S // 3].numpy(), g1))
          # ---- the same, starting from what is on disk: the BGZF bytes of the .vcf.gz in pinned host memory.  They cross
          # PCIe compressed, are inflated and parsed on the GPU; the matrix + site columns come back to pinned host memory.
          if not args.no_bgzf and rank == 0:
              import numpy as _np
              hdr = capi.synth_header(spec)
              full = _np.empty(len(hdr) + T, _np.uint8)
              full[:len(hdr)] = _np.frombuffer(hdr, _np.uint8)
              full[len(hdr):] = host.numpy()
              t0 = time.perf_counter()
              bg = capi.bgzf_compress_host(full, 6)
              t_comp = time.perf_counter() - t0
              del full
              bgp = torch.empty(bg.size, dtype=torch.uint8).pin_memory()
              bgp.numpy()[:] = bg
              del bg

              def bgzf_step():
                  q = capi.Parse.from_vcf_bytes(bgp.data_ptr(), region="chr22", device=local, nbytes=bgp.numel())
                  capi.check(capi.lib().hb_parse_fetch_matrix(q._h, out0.data_ptr(), out1.data_ptr()))
                  capi.check(capi.lib().hb_parse_fetch_sites(q._h, *[a.data_ptr() for a in sites]))
                  ms_inf = q.info.ms_inflate
                  q.close()
                  return ms_inf

              bgzf_step()
              torch.cuda.synchronize()
              t0 = time.perf_counter()
              ms_inf = [bgzf_step() for _ in range(args.e2e_steps)]
              torch.cuda.synchronize()
              dtb = time.perf_counter() - t0
              g0, g1 = p.sample(S // 3)
              e2e["from_bgzf_whole_file"] = {"value": float(V) * S / (dtb / args.e2e_steps), "unit": "calls/s", "h2d_bytes_per_step": int(bgp.numel()),
                                  "d2h_bytes_per_step": 2 * S * Vk + 10 * Vk, "steps": args.e2e_steps,
                                  "inflate_kernel_ms": sorted(ms_inf)[len(ms_inf) // 2], "text_over_bgzf": T / float(bgp.numel()),
                                  "matches_device_path": bool(np.array_equal(out0[S // 3].numpy(), g0) and np.array_equal(out1[S // 3].numpy(), g1)),
                                  "api": "hb_parse_vcf_bytes (BGZF in pinned host memory -> GPU inflate -> GPU parse) + hb_parse_fetch_matrix "
                                         "+ hb_parse_fetch_sites; rank 0 only", "bgzip_equivalent_host_s": t_comp}
              # streamed: slabs of BGZF members, H2D + inflate / parse / D2H overlapped (hb_parse_stream_bgzf_host)
              def bgzf_stream_step():
                  capi.check(capi.lib().hb_parse_stream_bgzf_host(bgp.data_ptr(), bgp.numel(), b"chr22", 1, local, args.slab_bytes,
                                                                  out0.data_ptr(), out1.data_ptr(), Vk, sites[0].data_ptr(),
                                                                  sites[1].data_ptr(), sites[2].data_ptr(), sites[3].data_ptr(),
                                                                  None, None, C.byref(nrec), None))
                  assert nrec.value == Vk

              out0.zero_(); out1.zero_()
              bgzf_stream_step()
              ok_stream = bool(np.array_equal(out0[S // 3].numpy(), g0) and np.array_equal(out1[S // 3].numpy(), g1))
              t0 = time.perf_counter()
              for _ in range(args.e2e_steps):
                  bgzf_stream_step()
              dts = time.perf_counter() - t0
              e2e["from_bgzf"] = {"value": float(V) * S / (dts / args.e2e_steps), "unit": "calls/s",
                                           "h2d_bytes_per_step": int(bgp.numel()), "d2h_bytes_per_step": 2 * S * Vk + 10 * Vk,
                                           "steps": args.e2e_steps, "matches_device_path": ok_stream,
                                           "api": "hb_parse_stream_bgzf_host: BGZF in pinned host memory -> slabs of %d MiB of text: "
                                                  "H2D compressed + GPU inflate / GPU parse / D2H overlapped; rank 0 only" % (args.slab_bytes >> 20)}
              del bgp
          del host, out0, out1
    except Exception as ex:          # e.g. not enough pinnable host memory on a crowded box: the device-resident numbers stand
        e2e = {"value": None, "unit": "calls/s", "error": "%s: %s" % (type(ex).__name__, ex)}

    if rank == 0:
        peak, peak_src = measured_peak()
        med = lambda a: float(sorted(a)[len(a) // 2])
        alg_tok = T
        alg_dec = 4.0 * Vk * S + 2.0 * Vk * S
        alg_parse = T + 2.0 * Vk * S + 33.0 * Vk
        step_s = ms / 1e3 / args.steps
        stages = {"locate_records": {"ms": med(tok), "gbs": alg_tok / (med(tok) / 1e3) / 1e9, "bytes": alg_tok,
                                     "note": "head walker reads ~2% of the text; tokenizer reads all of it"},
                  "sites": {"ms": med(sit)},
                  "decode_gt": {"ms": med(dec), "gbs": alg_dec / (med(dec) / 1e3) / 1e9, "bytes": alg_dec}}
        alg_all = alg_parse
        store = None
        if frames[0] is not None:
            fi = frames[0].info
            c_out = float(fi.total_bytes)
            alg_store = 2.0 * Vk * S + 33.0 * Vk + c_out
            alg_all = alg_parse + alg_store
            stages["site_templates"] = {"ms": med(k4["site_templates"])}
            alg_df = 2.0 * Vk * S + c_out           # allele planes read once, every frame written once
            stages["donor_frames"] = {"ms": med(k4["donor_frames"]), "bytes": alg_df,
                                      "gbs": alg_df / (med(k4["donor_frames"]) / 1e3) / 1e9,
                                      "note": "fused: allele-plane LZ4 + frame assembly into closed-form slots"}
            store = {"c_out_bytes": int(c_out), "frames": int(fi.n_chunks) * S, "chunk_records": int(fi.chunk_records),
                     "compression_ratio": float(fi.raw_bytes) / max(1.0, c_out), "raw_bytes_logical": int(fi.raw_bytes)}
        # the dominant kernel = the stage with the largest measured time that has algorithmic bytes
        dom = max((k for k in stages if "bytes" in stages[k]), key=lambda k: stages[k]["ms"])
        traffic, traffic_note = None, None
        try:        # DRAM bytes per launch: the ratio ncu measured for this kernel (profiles/) times this launch's algorithmic bytes
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01e_traffic.json")))[dom]
            traffic = tr["ratio"] * stages[dom]["bytes"]
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum = %.3f x algorithmic bytes in the ncu --set full capture "
                            "(200000 variants x 2504 samples, profiles/r01e_traffic.json), scaled to this launch" % tr["ratio"])
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": stages[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": stages[dom]["gbs"] / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src, "nominal_peak": NOMINAL_HBM_GBS,
                "algorithmic_bytes_per_launch": stages[dom]["bytes"], "ms_per_launch": stages[dom]["ms"],
                "path": {"algorithmic_bytes": alg_all, "gbs": alg_all / step_s / 1e9, "frac": alg_all / step_s / 1e9 / peak,
                         "definition": ("T + 2*(2*V'*S + 33*V') + C_out per step (SURVEY.md 8d 'fused total'), whole step incl. host syncs"
                                        if frames[0] is not None else
                                        "T + 2*V'*S + 33*V' per step (SURVEY.md 8d parse+decode), whole step incl. host syncs")},
                "stages": stages, "store": store}
        cpu = None
        if not args.no_cpu:
            nv = args.cpu_sample_variants
            cs = capi.synth_spec(nv, S, seed=args.seed)
            ctext = capi.synth_header(cs) + capi.synth_host(cs)
            cores = os.cpu_count() or 1
            v, dt, nd = cpu_baseline(ctext, capi.synth_sample_names(cs), nv, cores)
            cpu = {"value": v, "unit": "calls/s", "cores": cores, "kind": "port",
                   "sample": f"first {nv} variants x {S} samples ({len(ctext) / 1e6:.0f} MB); {nd} donors in parallel, one "
                             f"whole-text load_vcf pass per donor (reference driving pattern); {dt:.1f} s"}
        line = {"metric": "genotype calls/sec", "value": value, "unit": "calls/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": workload_config(args),
                "variants_per_s": float(V) * world / step_s, "records_kept": Vk, "text_bytes_per_rank": T,
                "step": "parse only (kernels 1-3)" if args.parse_only else "text -> Blosc2 frames (kernels 1-4)",
                "tokenizer": int(info.tokenizer_used), "parity_spot_check": parity,
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if frames[0] is not None:
        p.attach(None)
        frames[0].close()
    p.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
