#!/usr/bin/env python
"""bench.py -- headline benchmark of the composition / coverage scoring paths on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mags M] [--samples S] [--impl reference]

One "step" = one pass of the hot path over one batch of synthetic MAGs that is already resident in
HBM: canonical 4-mer + GC profiling of every contig (+ coverage of every (contig, sample) and the
per-MAG scores once those kernels are present).  Workload at N=1: BASELINE.json config 3
(1,000 synthetic MAGs, ~3 Gbp, 10 samples, one B200).  With N>1 every rank owns its own 1,000
MAGs (MAG-sharded, weak scaling, no data-path collective; a final all_gather of the per-contig
scores).  One JSON line on stdout (rank 0).

--impl reference times the CPU restatement of the reference algorithm (oracle/, the Rust crate
cannot be built here) on the box's host cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "contig Gbp/s k-mer profiling"
UNIT = "Gbp/s"


def ncu_traffic(kernel):
    """(DRAM bytes per launch of `kernel` from the committed ncu --set full capture, the file they came from), or
    (None, None).  The bench does not measure DRAM traffic itself (that needs ncu's replay)."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            return (json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]["dram_bytes"],
                    f"profiles/{name} (ncu --set full capture of this workload, not measured in this run)")
        except Exception:
            continue
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v == "Active"})
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(sm)}
        gpus = sorted({r[0] for r in self.rows if len(r) >= 9})
        if len(gpus) > 1:  # one sampler watching several GPUs: median SM clock and power of each
            out["sm_mhz_by_gpu"] = {g: float(np.median([float(r[1]) for r in self.rows if len(r) >= 9 and r[0] == g]))
                                    for g in gpus}
            out["power_w_by_gpu"] = {g: float(np.median([float(r[3]) for r in self.rows if len(r) >= 9 and r[0] == g
                                                         and r[3].replace(".", "").isdigit()] or [0.0])) for g in gpus}
        return out


# ---------------------------------------------------------------------------------------------

def cpu_tnf_sample(oracle, bases, offsets, mag_off, n_mags, threads):
    """Oracle (reference algorithm, C restatement) on the first n_mags MAGs; returns
    (seconds, bases processed, counts u32, gc f32)."""
    c1 = int(mag_off[n_mags])
    b1 = int(offsets[c1])
    sub_b = np.ascontiguousarray(bases[:b1])
    sub_o = np.ascontiguousarray(offsets[:c1 + 1])
    t0 = time.perf_counter()
    counts = oracle.get_tnf_concat(sub_b, sub_o, relative=False, threads=threads)
    t1 = time.perf_counter()
    gc = oracle.get_gc_concat(sub_b, sub_o, passes5=True)  # 5 passes, serial: as the reference does
    dt = time.perf_counter() - t0
    LAST_SPLIT.update(tnf_s=t1 - t0, gc_5pass_s=dt - (t1 - t0))
    return dt, b1, counts, gc


LAST_SPLIT = {}  # time split of the last cpu_tnf_sample call (tnf over contigs / serial 5-pass GC)


def fair_variant(oracle, bases, offsets, mag_off, n_mags):
    """the same CPU work with a ONE-pass GC count (the reference makes five passes over each
    sequence); reported next to the as-written figure, never instead of it."""
    c1 = int(mag_off[n_mags])
    b1 = int(offsets[c1])
    sub_b, sub_o = np.ascontiguousarray(bases[:b1]), np.ascontiguousarray(offsets[:c1 + 1])
    t0 = time.perf_counter()
    oracle.get_gc_concat(sub_b, sub_o, passes5=False)
    gc1 = time.perf_counter() - t0
    return {"value": b1 / (LAST_SPLIT["tnf_s"] + gc1) / 1e9, "unit": UNIT,
            "note": "get_tnf as written + one-pass GC instead of the reference's five passes",
            "tnf_only_gbp_s": b1 / LAST_SPLIT["tnf_s"] / 1e9, "gc_5pass_gbp_s": b1 / LAST_SPLIT["gc_5pass_s"] / 1e9,
            "gc_1pass_gbp_s": b1 / gc1 / 1e9}


def coverage_phase(args, lib, dev, d, d_off, world, rank, local, barrier):
    """Second metric of BASELINE.json: aligned reads/s of the coverage path.  One step = all S
    samples of this rank's MAGs: events (start, end) resident in HBM -> float32[n_contigs, S]."""
    import ctypes as C

    import torch
    import torch.distributed as dist

    from magpurify2_b200 import _coverage, _lib, synth

    lengths, mag_off, offsets = d["lengths"], d["mag_off"], d["offsets"]
    n, L, S = len(lengths), int(offsets[-1]), args.samples
    stream = torch.cuda.current_stream().cuda_stream
    # distinct event sets resident in HBM: all S of them when they fit (C3: 10 x 4.7 GB); a C4 shard (20 samples x
    # up to tens of GB) keeps as many as fit and the remaining columns re-use them in turn -- every column is still
    # computed from its events by the same kernels, only the data repeats
    resident = []
    for s in range(S):
        if args.resident_samples and len(resident) >= args.resident_samples:
            break
        if resident:
            free, _tot = torch.cuda.mem_get_info(dev)
            need = 8 * resident[0][3]
            if free < 4.5 * need + (6 << 30):  # the generator's temporaries + the kernels' scratch
                break
        resident.append(synth.make_events_device(lengths, mag_off, s, dev, seed=args.seed + 1000 * rank,
                                                 depth_scale=args.depth_scale))
    # span / disorder of every resident set, measured once by the library's own order kernels -- what the BAM decoder
    # hands over with its events (mp2_bam_max_block / mp2_bam_max_disorder); --max-span 0 measures inside every call
    order = []
    for st_, en_, eo_, R_ in resident:
        sp, ds = C.c_uint32(0), C.c_uint32(0)
        _lib.check(lib.mp2_cov_order_device(st_.data_ptr(), en_.data_ptr(), eo_.data_ptr(), d_off.data_ptr(), n, R_,
                                            C.byref(sp), C.byref(ds), stream))
        order.append((int(sp.value), int(ds.value)))
    samples = [resident[s % len(resident)] + order[s % len(resident)] for s in range(S)]
    R_total = sum(smp[3] for smp in samples)
    cov = torch.empty((n, S), dtype=torch.float32, device=dev)
    n_mags = len(mag_off) - 1
    d_len = torch.from_numpy(lengths.astype(np.int64)).to(dev)
    d_mag_off = torch.from_numpy(mag_off.astype(np.int64)).to(dev)
    d_sel = torch.empty((n_mags, S), dtype=torch.uint8, device=dev)
    d_score = torch.empty(n, dtype=torch.float64, device=dev)

    # samples are independent (one column of `cov` each) and may alternate between streams (--cov-streams) so
    # that the prologue kernels of sample s+1 run under the tail of sample s; measured on C3 that is 15 % faster
    # at best and 50 % slower at worst (run-to-run), so the default keeps them on one stream
    main_stream = torch.cuda.current_stream()
    lanes = (main_stream,) + tuple(torch.cuda.Stream(device=dev) for _ in range(max(1, args.cov_streams) - 1))

    def step(lanes=lanes, which=None, span=None, disorder=0, score=True):
        which = samples if which is None else which
        if span is None and args.max_span >= 0:
            span = args.max_span  # -1 (default): vouch with the setup measurement; 0: measure in every call
        for ls in lanes[1:]:
            ls.wait_stream(main_stream)
        for s, (st, en, eo, R, sp_m, ds_m) in enumerate(which):
            sp_, ds_ = (sp_m, ds_m) if span is None else (span, disorder)  # None: what was measured at setup
            _lib.check(lib.mp2_cov_events_device(st.data_ptr(), en.data_ptr(), eo.data_ptr(), d_off.data_ptr(),
                                                 n, R, L, sp_, ds_, 75, args.trim, args.trim,
                                                 cov.data_ptr() + 4 * s, S, None, lanes[s % len(lanes)].cuda_stream))
        for ls in lanes[1:]:
            main_stream.wait_stream(ls)
        if not score:
            return
        # core.Coverage: samples with length-weighted mean coverage >= 1, then base-25 log-ratio score
        _lib.check(lib.mp2_select_samples_device(cov.data_ptr(), d_len.data_ptr(), d_mag_off.data_ptr(), n_mags, n, S,
                                                 1.0, d_sel.data_ptr(), None, stream))
        _lib.check(lib.mp2_logratio_device(cov.data_ptr(), d_len.data_ptr(), d_mag_off.data_ptr(), n_mags, n, S,
                                           d_sel.data_ptr(), 25.0, 1e-5, 0.0, 1.0, 3, 0, d_score.data_ptr(), None,
                                           stream))

    def timed(**kw):
        """ms per step of `steps` steps (device events on the launching stream), after one untimed step."""
        step(**kw)
        torch.cuda.synchronize()
        a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a_.record()
        for _ in range(args.steps):
            step(**kw)
        b_.record()
        torch.cuda.synchronize()
        return a_.elapsed_time(b_) / args.steps

    def path_of(sample, span, disorder):
        """Which path computed one sample: share of the contigs (with a window) the general path wrote."""
        st, en, eo, R = sample[:4]
        d_stats = torch.zeros(n * 56, dtype=torch.uint8, device=dev)
        _lib.check(lib.mp2_cov_events_device(st.data_ptr(), en.data_ptr(), eo.data_ptr(), d_off.data_ptr(), n, R, L,
                                             span, disorder, 75, args.trim, args.trim, cov.data_ptr(), S,
                                             d_stats.data_ptr(), stream))
        torch.cuda.synchronize()
        fl = d_stats.cpu().numpy().view(_lib.COV_STATS_DTYPE)["flags"]
        return "general (difference array in HBM)" if (fl & 1).any() else "shared-memory streaming kernel"

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = lib.mp2_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cov_sampler = ClockSampler(",".join(str(i) for i in range(world))) if rank == 0 else None
    if cov_sampler:
        cov_sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    cov_clocks = cov_sampler.stop() if cov_sampler else None
    launches = lib.mp2_launch_count() - l0
    # per-kernel durations for the roofline: the same steps on ONE stream (two samples in flight overlap on
    # the SMs, so a launch's event-to-event time would include the other sample's work)
    lib.mp2_profile_reset()
    lib.mp2_profile_enable(1)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        step(lanes=(main_stream,))
    e3.record()
    barrier()
    ms_one_stream = e2.elapsed_time(e3)
    kern = {}
    for name in ("cov_stream2_kernel", "cov_stream_kernel", "cov_tile_kernel", "cov_scatter_kernel", "cov_depth_kernel",
                 "cov_deep_kernel", "cov_order_kernel", "cov_order_scan_kernel", "cov_bounds_kernel",
                 "cov_zero_kernel"):
        k_ms, k_n = C.c_double(0), C.c_int64(0)
        lib.mp2_profile_read(name.encode(), C.byref(k_ms), C.byref(k_n))
        kern[name] = (k_ms.value, k_n.value)
    lib.mp2_profile_enable(0)
    # ---- secondary legs (N = 1 only): the same samples with the order vouched for instead of measured, and
    # SURVEY 8(d)'s variant (read length N(150,15), 2 % of the reads with one I / D = two blocks, out of order)
    legs = {}
    if world == 1 and args.cov_legs:
        legs["headline_path"] = path_of(samples[0], samples[0][4], samples[0][5])
        legs["headline_order"] = {"max_span": samples[0][4], "max_disorder": samples[0][5]}
        legs["order_vouched_ms_per_sample"] = timed(score=False, span=None if args.max_span < 0 else 150) / S
        legs["order_measured_in_call_ms_per_sample"] = timed(span=0, disorder=0, score=False) / S
        nv = max(1, min(S, args.variant_samples))
        var = []
        for s in range(nv):
            v = synth.make_events_device(lengths, mag_off, s, dev, seed=args.seed + 1000 * rank,
                                         depth_scale=args.depth_scale, variant=True)
            sp, ds = C.c_uint32(0), C.c_uint32(0)
            _lib.check(lib.mp2_cov_order_device(v[0].data_ptr(), v[1].data_ptr(), v[2].data_ptr(), d_off.data_ptr(), n,
                                                v[3], C.byref(sp), C.byref(ds), stream))
            var.append(tuple(v) + (int(sp.value), int(ds.value)))
        Rv = sum(v[3] for v in var)
        t_meas = timed(which=var, span=0, disorder=0, score=False) / nv
        save, args.max_span = args.max_span, -1
        t_vouch = timed(which=var, score=False) / nv
        args.max_span = save
        legs["variant"] = {
            "workload": "read length N(150,15) clipped to [50,250], 2 % of the reads with one 1-3 bp I or D (two "
                        "blocks; the second one starts after the next reads' first blocks)",
            "samples": nv, "blocks_per_sample": Rv / nv,
            "measured_order": {"max_span": var[0][4], "max_disorder": var[0][5]},
            "ms_per_sample_order_measured_in_call": t_meas, "ms_per_sample_order_vouched": t_vouch,
            "reads_per_s_order_vouched": Rv / nv / (t_vouch * 1e-3),
            "path": path_of(var[0], var[0][4], var[0][5]),
        }
        del var
        step()  # leave `cov` holding the headline samples for the checks below
        torch.cuda.synchronize()
    ms_ranks = [ms]
    if world > 1:
        g = torch.zeros(world, 3, dtype=torch.float64, device=dev)
        mine = torch.tensor([ms, ms_one_stream, float(R_total)], dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g, mine)
        ms_ranks = [round(float(x), 3) for x in g[:, 0]]
        ms_ranks_one = [round(float(x), 3) for x in g[:, 1]]
        ms = float(g[:, 0].max().item())
        R_all = float(g[:, 2].sum().item())
    else:
        R_all = float(R_total)
        ms_ranks_one = [ms_one_stream]
    value = R_all * args.steps / (ms * 1e-3)

    # end to end: one sample through the host-buffer API (pinned events in, coverage column out)
    st, en, eo, R0 = samples[0][:4]
    h_st = torch.empty(R0, dtype=torch.int32, pin_memory=True); h_st.copy_(st)
    h_en = torch.empty(R0, dtype=torch.int32, pin_memory=True); h_en.copy_(en)
    torch.cuda.synchronize()
    h_eo = eo.cpu().numpy()
    a_st, a_en = h_st.numpy().view(np.uint32), h_en.numpy().view(np.uint32)
    vouch = dict(max_span=samples[0][4], max_disorder=samples[0][5])  # as get_bam_coverages passes what the reader measured
    _coverage.cov_from_events(a_st, a_en, h_eo, offsets, 75, args.trim, args.trim, device=local, **vouch)
    barrier()
    t0 = time.perf_counter()
    h_cov = _coverage.cov_from_events(a_st, a_en, h_eo, offsets, 75, args.trim, args.trim, device=local, **vouch)
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, float(R0)], dtype=torch.float64, device=dev)
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        e2e_val = float(t[1].item()) / float(tm[0].item())
    else:
        e2e_val = R0 / e2e_s
    assert np.array_equal(h_cov, cov[:, 0].cpu().numpy(), equal_nan=True)
    if rank != 0:
        return None
    peak, peak_src = peaks()
    per_launch = lambda k: kern[k][0] / max(1, kern[k][1])
    R_avg = R_total / S
    # Roofline of the coverage kernel.  `achieved` follows SURVEY.md §8(d): algorithmic bytes of one sample
    # (8 B/base for the int32 difference array written and read once + 20 B/read for the event and its two
    # updates) over the kernel time.  cov_stream_kernel keeps the difference array in shared memory, so the
    # bytes it actually has to move (`moved_*`: events once + bounds + results) are ~7x fewer; both are given.
    sample_bytes = 8.0 * L + 20.0 * R_avg
    moved = {"cov_stream_kernel": 8.0 * R_avg + 24.0 * (L / 65536.0) + 4.0 * n,
             "cov_tile_kernel": 8.0 * R_avg + 24.0 * (L / 16384.0) + 4.0 * n,
             "cov_depth_kernel": 4.0 * (L + 1), "cov_scatter_kernel": 8.0 * R_avg + 8.0 * R_avg}
    moved["cov_stream2_kernel"] = moved["cov_stream_kernel"]
    sample_ms = ms_one_stream / args.steps / S  # per sample when the samples run one after the other
    rl = []
    for name in ("cov_stream2_kernel", "cov_stream_kernel", "cov_tile_kernel", "cov_depth_kernel", "cov_scatter_kernel"):
        t_ms = per_launch(name)
        if t_ms < 0.02 * sample_ms:  # not launched, or gated off (exits at once): not on the path of this run
            continue
        fused = name in ("cov_stream2_kernel", "cov_stream_kernel", "cov_tile_kernel")  # one kernel does the whole sample
        bytes_ = sample_bytes if fused else moved[name]
        ach = bytes_ / (t_ms * 1e-3) / 1e9
        r = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
             "traffic": ncu_traffic(name)[0] if (args.mags == 1000 and args.depth_scale == 1.0) else None,
             "traffic_source": ncu_traffic(name)[1],
             "algorithmic_bytes_per_launch": bytes_, "kernel_ms": t_ms, "launches_timed": kern[name][1]}
        if fused:
            r["algorithmic_bytes"] = "SURVEY 8(d): 8*L + 20*R per sample"
            r["moved_bytes_per_launch"] = moved[name]
            r["moved_gbs"] = moved[name] / (t_ms * 1e-3) / 1e9
            r["moved_frac_of_peak"] = r["moved_gbs"] / peak
        rl.append(r)
    dom = max(rl, key=lambda r: r["kernel_ms"])
    rec = {
        "metric": "aligned reads/s coverage", "value": value, "unit": "reads/s", "ms_per_step": ms / args.steps,
        "config": {"workload": f"{args.config.upper()}: coverage of {args.mags} MAGs x {S} samples on this GPU, {R_total / S / 1e6:.0f} M aligned blocks/sample",
                   "samples": S, "resident_event_sets": len(resident), "reads_per_step": R_total, "positions": L,
                   "trim": args.trim, "end_exclusion": 75,
                   "input": "(start,end) uint32 pairs grouped by contig, resident in HBM",
                   "order": ("every call vouched for with the block span / start disorder mp2_cov_order_device measured for "
                             "that sample at set-up, outside the timed region (the product path: mp2_bam_open hands the "
                             "same two numbers over with the events it decodes)" if args.max_span < 0 else
                             "block span and start disorder measured on the device inside every call (timed)"
                             if args.max_span == 0 else f"vouched by the caller: span <= {args.max_span}, sorted"),
                   "streams": len(lanes),
                   "note": "kernel_ms / roofline are from a second pass of the same steps on one stream"},
        "ms_per_step_one_stream": ms_one_stream / args.steps,
        "clocks": cov_clocks,
        "ms_per_step_by_rank": [round(x / args.steps, 3) for x in ms_ranks],
        "ms_per_step_by_rank_second_pass": [round(x / args.steps, 3) for x in ms_ranks_one],
        "roofline": dict(dom, peak_source=peak_src), "kernels": rl,
        "sample_bytes_algorithmic": sample_bytes,
        "sample_gbs": sample_bytes / (sample_ms * 1e-3) / 1e9,
        "e2e": {"value": e2e_val, "unit": "reads/s", "h2d_bytes_per_step": 8 * R0 + 16 * (n + 1),
                "d2h_bytes_per_step": 4 * n, "steps": 1,
                "api": "magpurify2_b200._coverage.cov_from_events (mp2_cov_events, pinned host buffers, 1 sample)"},
        "gpu_launches": int(launches),
        "deep_fallback_ms": kern["cov_deep_kernel"][0] / max(1, kern["cov_deep_kernel"][1]),
        "order_check_ms_per_sample": sum(kern[k][0] for k in ("cov_order_kernel", "cov_order_scan_kernel"))
                                     / max(1, args.steps * S),
        "kernel_ms_per_sample": {k: round(v[0] / max(1, args.steps * S), 4) for k, v in kern.items() if v[1]},
        "legs": legs,
    }
    if not args.no_cpu:
        import oracle

        m = min(args.mags, max(1, args.ref_mags))
        c1 = int(mag_off[m])
        e1_ = int(h_eo[c1])
        t0 = time.perf_counter()
        c_cov, _ = oracle.cov_sample(a_st[:e1_], a_en[:e1_], h_eo[:c1 + 1], lengths[:c1], 75, args.trim, args.trim, threads=1)
        dt = time.perf_counter() - t0
        ok = bool(np.array_equal(c_cov, h_cov[:c1], equal_nan=True))
        # scores of the first MAG against the oracle (all samples)
        m0 = slice(0, int(mag_off[1]))
        cov0 = cov[m0].cpu().numpy()
        sel0, _ = oracle.select_samples(cov0, lengths[m0], 1.0)
        if sel0.any() and cov0.shape[0] > 2:
            want0, _ = oracle.log_ratio_scores(cov0[:, sel0], lengths[m0], 25)
            ok = ok and bool(np.allclose(d_score[m0].cpu().numpy(), want0, atol=1e-6))
        rec["cpu_baseline"] = {"value": e1_ / dt, "unit": "reads/s", "cores": 1, "kind": "port",
                               "sample": f"sample 0, first {m} MAGs ({e1_ / 1e6:.1f} M blocks, {int(offsets[c1]) / 1e6:.0f} Mbp); "
                                         "single thread per sample as coverm's record loop is (events already decoded)",
                               "parity_with_gpu": ok}
        if not ok:
            raise SystemExit("PARITY FAILURE: GPU coverage differs from the CPU oracle")
    return rec


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    from magpurify2_b200 import synth

    oracle.build()
    threads = os.cpu_count() or 1
    n_mags = min(args.mags, args.ref_mags)
    d = synth.make_mags(n_mags, seed=args.seed, decorate=True)
    L = int(d["offsets"][-1])
    times = []
    for it in range(args.warmup + args.steps):
        dt, _b, _c, _g = cpu_tnf_sample(oracle, d["bases"], d["offsets"], d["mag_off"], n_mags, threads)
        if it >= args.warmup:
            times.append(dt)
    t = float(np.sum(times))
    val = L * len(times) / t / 1e9
    fair = fair_variant(oracle, d["bases"], d["offsets"], d["mag_off"], n_mags)
    sample = f"{n_mags} of {args.mags} MAGs per step ({L / 1e6:.1f} Mbp): get_tnf with {threads} threads over contigs + serial 5-pass get_gc"
    # coverage leg of the reference arm: coverm's record loop is single-threaded per BAM and the BAMs
    # are processed one after the other; events are already decoded here
    cov_rec = None
    if args.samples > 0:
        m = max(1, n_mags // 2)
        c1 = int(d["mag_off"][m])
        st, en, eo = synth.make_events(d["lengths"][:c1], d["mag_off"][:m + 1], 0, seed=args.seed,
                                       depth_scale=args.depth_scale)
        ctimes = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle.cov_sample(st, en, eo, d["lengths"][:c1], 75, args.trim, args.trim, threads=1)
            if it >= args.warmup:
                ctimes.append(time.perf_counter() - t0)
        cval = len(st) * len(ctimes) / float(np.sum(ctimes))
        cs = (f"1 sample x first {m} MAGs per step ({len(st) / 1e6:.1f} M aligned blocks, "
              f"{int(d['offsets'][c1]) / 1e6:.0f} Mbp), single thread as coverm's record loop")
        cov_rec = {"metric": "aligned reads/s coverage", "value": cval, "unit": "reads/s",
                   "ms_per_step": float(np.mean(ctimes)) * 1e3,
                   "cpu_baseline": {"value": cval, "unit": "reads/s", "cores": 1, "kind": "port", "sample": cs},
                   "e2e": {"value": cval, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / len(times) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"C3: composition on {args.mags} synthetic MAGs (~3 Mbp, ~300 contigs each)",
                   "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "fair_variant": fair},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if cov_rec:
        line["coverage"] = cov_rec
    emit(line)
    return 0


_JSON_OUT = None


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter
    written to fd 1) was redirected to stderr at start-up."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mags", type=int, default=1000, help="MAGs per GPU (C3: 1000)")
    ap.add_argument("--samples", type=int, default=None, help="samples per step (C3: 10, C4: 20)")
    ap.add_argument("--config", default="c3", choices=["c3", "c4"],
                    help="c3 (default): every GPU its own 1,000 MAGs x 10 samples, weak scaling; c4: ONE set of 10,000 "
                         "MAGs x 20 samples LPT-sharded over the GPUs, strong scaling (2,500 MAGs on a single GPU)")
    ap.add_argument("--c4-mags", type=int, default=0, help="MAGs of the C4 set (default 10000; 2500 on one GPU)")
    ap.add_argument("--resident-samples", type=int, default=0,
                    help="distinct event sets kept in HBM (default: as many as fit); the other columns reuse them")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-mags", type=int, default=300, help="MAGs per step of the CPU arms")
    ap.add_argument("--long-contigs", action="store_true", help="C5 shape instead of C3")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--trim", type=float, default=0.05)
    ap.add_argument("--max-span", type=int, default=-1,
                    help="-1 (default) = every call is vouched for with the span / disorder the library measured for that "
                         "sample at set-up (as the BAM decoder does for its own output); 0 = measured inside every call, "
                         "i.e. inside the timed region; > 0 = this span, sorted starts")
    ap.add_argument("--no-cov-legs", dest="cov_legs", action="store_false",
                    help="skip the secondary coverage legs (vouched order, indel variant)")
    ap.add_argument("--variant-samples", type=int, default=2)
    ap.add_argument("--depth-scale", type=float, default=1.0)
    ap.add_argument("--cov-streams", type=int, default=1,
                    help="streams the coverage samples alternate between (2: 51..91 ms per C3 step, unstable; 1: 59.9 ms)")
    args = ap.parse_args()
    args.samples_given = args.samples is not None
    if args.samples is None:
        args.samples = 10
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from magpurify2_b200 import _composition, _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: magpurify2_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from magpurify2_b200 import numa

    placement = numa.bind_to_gpu_node(local) if world > 1 else {"device": local, "bound": False, "why": "single GPU"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # ---- synthetic shard of this rank (MAGs are independent: shard = MAG set) ----
    c4 = None
    if args.config == "c4":
        # BASELINE config 4: ONE set of MAGs (10 000, ~30 Gbp) x 20 samples, whole MAGs dealt out to the ranks by
        # longest-processing-time-first on their base pairs (sharding.lpt_partition): strong scaling
        from magpurify2_b200 import sharding

        total_mags = args.c4_mags if args.c4_mags else (10000 if world > 1 else 2500)
        args.samples = args.samples if args.samples_given else 20
        clen_all, mag_off_all, gc_all = synth.mag_layout(total_mags, seed=args.seed)
        bp_all = np.add.reduceat(clen_all, mag_off_all[:-1])
        parts = sharding.lpt_partition(bp_all, world)
        mine = parts[rank]
        c4 = {"total_mags": int(total_mags), "total_bp": int(bp_all.sum()), "mags_of_rank": [int(len(p)) for p in parts],
              "bp_of_rank": [int(bp_all[p].sum()) for p in parts]}
        args.mags = int(len(mine))
        d = synth.make_mags_device(len(mine), dev, seed=args.seed + 1000 * rank,
                                   layout=synth.subset_layout(clen_all, mag_off_all, gc_all, mine))
    else:
        d = synth.make_mags_device(args.mags, dev, seed=args.seed + 1000 * rank, long_contigs=args.long_contigs)
    bases, offsets, mag_off = d["bases"], d["offsets"], d["mag_off"]
    n = len(offsets) - 1
    L = int(offsets[-1])
    d_off = torch.from_numpy(offsets).to(dev)
    d_counts = torch.empty((n, 136), dtype=torch.int32, device=dev)
    d_rel = torch.empty((n, 136), dtype=torch.float32, device=dev)
    d_gc = torch.empty(n, dtype=torch.float32, device=dev)
    d_len = torch.from_numpy(d["lengths"].astype(np.int64)).to(dev)
    d_mag_off = torch.from_numpy(mag_off.astype(np.int64)).to(dev)
    n_mags = len(mag_off) - 1
    d_dist = torch.empty(n, dtype=torch.float64, device=dev)
    d_scores = torch.empty((2, n), dtype=torch.float64, device=dev)  # tnf_score, gc_content_score
    stream = torch.cuda.current_stream().cuda_stream
    if world > 1:
        nmax = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(nmax, op=dist.ReduceOp.MAX)
        n_pad = int(nmax.item())
        g_in = torch.zeros((2, n_pad), dtype=torch.float64, device=dev)
        g_out = torch.empty((world, 2, n_pad), dtype=torch.float64, device=dev)

    main_s = torch.cuda.current_stream()
    side_s = torch.cuda.Stream(device=dev)  # the GC score does not depend on the CLR distances: it runs beside them

    def step():
        # counts + GC of every contig; the relative profiles (get_tnf's output) are normalised on the side stream
        _lib.check(lib.mp2_tnf_device(bases.data_ptr(), d_off.data_ptr(), n, L, d_counts.data_ptr(),
                                      None, d_gc.data_ptr(), stream))
        # per-MAG scores: relative TNF + GC log-ratio -> gc_content_score (side stream) beside
        # CLR/centroid distance -> tnf_score (neither needs the other)
        side_s.wait_stream(main_s)
        _lib.check(lib.mp2_tnf_normalize_device(d_counts.data_ptr(), n, d_rel.data_ptr(), side_s.cuda_stream))
        _lib.check(lib.mp2_logratio_device(d_gc.data_ptr(), d_len.data_ptr(), d_mag_off.data_ptr(), n_mags, n, 1, None,
                                           2.0, 1e-5, 0.0, 1.0, 3, 0, d_scores[1].data_ptr(), None, side_s.cuda_stream))
        _lib.check(lib.mp2_clr_centroid_device(d_counts.data_ptr(), d_len.data_ptr(), d_mag_off.data_ptr(), n_mags, n,
                                               None, d_dist.data_ptr(), stream))
        d32 = d_dist.to(torch.float32)
        _lib.check(lib.mp2_logratio_device(d32.data_ptr(), d_len.data_ptr(), d_mag_off.data_ptr(), n_mags, n, 1, None,
                                           4.0, 1e-5, 0.0, 1.0, 3, 1, d_scores[0].data_ptr(), None, stream))
        main_s.wait_stream(side_s)
        if world > 1:  # the only exchange: final gather of the per-contig scores (NCCL over NVLink).  It is
            # asynchronous: the gather of batch i runs under the kernels of batch i+1 (the last one is waited for
            # inside the timed region)
            if pending[0] is not None:
                pending[0].wait()  # g_in / g_out are free again
            g_in[:, :n] = d_scores
            pending[0] = dist.all_gather_into_tensor(g_out, g_in, async_op=True)

    pending = [None]

    def barrier():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    lib.mp2_profile_reset()
    lib.mp2_profile_enable(1)
    launches0 = lib.mp2_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    for i in range(args.steps):
        marks[i].record()
        step()
    if pending[0] is not None:  # the last batch's gather belongs to the timed region
        pending[0].wait()
        pending[0] = None
    marks[-1].record()
    barrier()
    e0, e1 = marks[0], marks[-1]
    ms = e0.elapsed_time(e1)
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.mp2_launch_count() - launches0
    k_ms, k_n = C.c_double(0), C.c_int64(0)
    lib.mp2_profile_read(b"tnf_kernel", C.byref(k_ms), C.byref(k_n))
    comp_kernels = {}
    for name in ("tnf_kernel", "tnf_chunk_map_kernel", "tnf_rel_kernel", "tnf_gc_kernel", "clr_mag_kernel",
                 "wmedian_kernel", "logratio_kernel"):
        a_ms, a_n = C.c_double(0), C.c_int64(0)
        lib.mp2_profile_read(name.encode(), C.byref(a_ms), C.byref(a_n))
        comp_kernels[name] = round(a_ms.value / args.steps, 4)
    lib.mp2_profile_enable(0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        tot = torch.tensor([float(L)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        L_all = float(tot.item())
    else:
        L_all = float(L)
    value = L_all * args.steps / (ms * 1e-3) / 1e9

    # ---- the same counts from the north-star 2-bit packed layout (codes + validity bits) ----
    packed = None
    try:
        nq = (L + 63) // 64
        d_codes = torch.empty(nq * 4, dtype=torch.int32, device=dev)
        d_valid = torch.empty(nq * 2, dtype=torch.int32, device=dev)
        d_counts2 = torch.empty((n, 136), dtype=torch.int32, device=dev)
        pe0, pe1, pe2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        for _ in range(2):
            _lib.check(lib.mp2_pack_device(bases.data_ptr(), L, d_codes.data_ptr(), d_valid.data_ptr(), stream))
            _lib.check(lib.mp2_tnf_packed_device(d_codes.data_ptr(), d_valid.data_ptr(), d_off.data_ptr(), n, L,
                                                 d_counts2.data_ptr(), None, None, stream))
        torch.cuda.synchronize()
        pe0.record()
        for _ in range(args.steps):
            _lib.check(lib.mp2_pack_device(bases.data_ptr(), L, d_codes.data_ptr(), d_valid.data_ptr(), stream))
        pe1.record()
        for _ in range(args.steps):
            _lib.check(lib.mp2_tnf_packed_device(d_codes.data_ptr(), d_valid.data_ptr(), d_off.data_ptr(), n, L,
                                                 d_counts2.data_ptr(), None, None, stream))
        pe2.record()
        torch.cuda.synchronize()
        same = bool(torch.equal(d_counts2, d_counts))
        t_pack, t_cnt = pe0.elapsed_time(pe1) / args.steps, pe1.elapsed_time(pe2) / args.steps
        pk_bytes = (L + 3) // 4 + (L + 7) // 8 + n * 560
        packed = {"pack_ms": t_pack, "pack_gbs": (L + pk_bytes - n * 560) / (t_pack * 1e-3) / 1e9,
                  "count_ms": t_cnt, "count_gbp_s": L / (t_cnt * 1e-3) / 1e9,
                  "count_algorithmic_gbs": pk_bytes / (t_cnt * 1e-3) / 1e9,
                  "counts_identical_to_ascii_path": same}
        del d_codes, d_valid, d_counts2
    except Exception as ex:  # noqa: BLE001 -- informative only
        packed = {"error": str(ex)}

    # ---- coverage phase: S samples of device-resident alignment blocks ----
    cov_rec = None
    if args.samples > 0:
        cov_rec = coverage_phase(args, lib, dev, d, d_off, world, rank, local, barrier)

    # ---- end to end through the public API with HOST buffers (H2D + D2H inside the timed region) ----
    h_bases = torch.empty(L, dtype=torch.uint8, pin_memory=True)
    h_bases.copy_(bases)
    torch.cuda.synchronize()
    hb = h_bases.numpy()
    e2e_steps = max(1, min(args.steps, 3))
    o_rel = _composition.pinned_empty((n, 136), np.float32)
    o_gc = torch.empty(n, dtype=torch.float32, pin_memory=True).numpy()
    _composition.tnf_from_concat(hb, offsets, relative=True, device=local, out_rel=o_rel, out_gc=o_gc)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _c, h_rel, h_gc = _composition.tnf_from_concat(hb, offsets, relative=True, device=local, out_rel=o_rel,
                                                       out_gc=o_gc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_ascii = L_all * e2e_steps / e2e_s / 1e9
    # the host-buffer path and the device-pointer path must agree bit for bit
    assert np.array_equal(h_gc.view(np.uint32), d_gc.cpu().numpy().view(np.uint32))
    assert np.array_equal(h_rel.view(np.uint32), d_rel.cpu().numpy().view(np.uint32))
    # ---- the product's upload path: the run's stream 2-bit packed on the host (once, when the genomes are read:
    # multi.composition_of -> pack_parts), then 0.375 B/base over PCIe.  Timed from the packed pinned buffers.
    host_threads = max(1, min(len(os.sched_getaffinity(0)), (os.cpu_count() or 1) // max(1, world)))
    t0 = time.perf_counter()
    codes, valid = _composition.pack_parts([hb], threads=host_threads)
    pack_first_s = time.perf_counter() - t0  # includes page-locking the two output buffers
    t0 = time.perf_counter()
    _lib.load().mp2_pack_host((C.c_void_p * 1)(hb.ctypes.data), _lib.ptr(np.array([L], dtype=np.int64)), 1,
                              _lib.ptr(codes), _lib.ptr(valid), host_threads)
    pack_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    runs = _composition.invalid_runs(valid, L, threads=host_threads)
    runs_s = time.perf_counter() - t0
    # (a) codes + validity mask, 0.375 B/base
    _composition.tnf_from_packed(codes, valid, offsets, relative=True, device=local, out_rel=o_rel, out_gc=o_gc)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        _composition.tnf_from_packed(codes, valid, offsets, relative=True, device=local, out_rel=o_rel, out_gc=o_gc)
    torch.cuda.synchronize()
    e2e_mask = L * e2e_steps / (time.perf_counter() - t0) / 1e9
    # (b) codes + runs of invalid bases, 0.25 B/base: what the composition driver uploads
    _composition.tnf_from_packed(codes, None, offsets, relative=True, device=local, out_rel=o_rel, out_gc=o_gc, runs=runs)
    barrier()
    e2e_times = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        _c, h_rel, h_gc = _composition.tnf_from_packed(codes, None, offsets, relative=True, device=local,
                                                       out_rel=o_rel, out_gc=o_gc, runs=runs)
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t0)
    e2e_p = float(np.sum(e2e_times))
    if world > 1:
        t = torch.tensor([e2e_p], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_p = float(t.item())
    e2e_val = L_all * e2e_steps / e2e_p / 1e9
    assert np.array_equal(h_gc.view(np.uint32), d_gc.cpu().numpy().view(np.uint32))
    assert np.array_equal(h_rel.view(np.uint32), d_rel.cpu().numpy().view(np.uint32))
    h2d = codes.nbytes + runs.nbytes + 8 * (n + 1)
    d2h = n * 136 * 4 + n * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    alg_bytes = L + n * (136 * 4 + 8 + 8)  # bases once + counts/gc rows once + offsets once
    k_avg_ms = k_ms.value / max(1, k_n.value)
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "ms_per_step_min_median_max": [round(float(f(step_ms)), 4) for f in (np.min, np.median, np.max)],
        "higher_is_better": True,
        "scaling": "strong" if c4 else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {
            "workload": ("C5 long-contig stress" if args.long_contigs else
                         f"C4: composition on ONE set of {c4['total_mags']} synthetic MAGs ({c4['total_bp'] / 1e9:.1f} Gbp), "
                         f"whole MAGs LPT-sharded over {world} GPU(s)" if c4 else
                         f"C3: composition on {args.mags} synthetic MAGs per GPU (~3 Mbp, ~300 contigs each)"),
            "c4": c4,
            "mags_per_gpu": args.mags, "contigs_per_gpu": n, "bases_per_gpu": L,
            "input": "ASCII, 1 B/base, resident in HBM", "l2": "inputs (GBs) far larger than the 126 MB L2",
            "parallelism": f"mag-shard x{world}",
            "step": "mp2_tnf_device (counts, GC) + CLR/centroid + weighted-median log-ratio score on the main stream; "
                    "mp2_tnf_normalize_device (relative TNF) + the GC log-ratio score on a side stream beside them"
                    + (" + all_gather of the scores" if world > 1 else ""),
        },
        "roofline": {"bound": "hbm", "kernel": "tnf_kernel", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic("tnf_kernel")[0] if (args.mags == 1000 and not args.long_contigs) else None,
                     "traffic_source": ncu_traffic("tnf_kernel")[1],
                     "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": k_avg_ms,
                     "launches_timed": k_n.value, "peak_source": peak_src},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step_min_median_max": [round(float(f(e2e_times)) * 1e3, 3) for f in (np.min, np.median, np.max)],
                "api": "magpurify2_b200._composition.tnf_from_packed(runs=, out_rel=, out_gc=) -> mp2_tnf_packed: 2-bit codes "
                       "in pinned host memory + the runs of invalid bases in (0.25 B/base over PCIe; the device rebuilds "
                       "the validity mask), relative TNF + GC in pinned host memory out, rows of finished contigs copied "
                       "back under the upload of the later slices",
                "invalid_runs": {"count": int(len(runs)), "extract_ms": runs_s * 1e3},
                "with_validity_mask_0.375_B_per_base": {"value": e2e_mask, "unit": UNIT},
                "input": "the run's stream as the ingest leaves it: packed on the host by mp2_pack_host from the FASTA "
                         "reader's buffers, once per run, OUTSIDE the timed region (see host_pack); covers counts, TNF "
                         "and GC, not the CLR / log-ratio scores and not FASTA parsing",
                "host_pack": {"gb_per_s": L / pack_s / 1e9, "threads": host_threads, "ms": pack_s * 1e3,
                              "first_call_ms_incl_page_locking": pack_first_s * 1e3},
                "ascii_upload": {"value": e2e_ascii, "unit": UNIT, "h2d_bytes_per_step": L + 8 * (n + 1),
                                 "api": "tnf_from_concat -> mp2_tnf, 1 B/base over PCIe (round 1's e2e)"},
                "from_ascii_incl_host_pack": {"value": L_all / ((e2e_p / e2e_steps) + pack_s) / 1e9, "unit": UNIT,
                                              "note": "pack, then upload, not overlapped"}},
        "kernel_ms_per_step": comp_kernels,
        "packed_2bit_path": packed,
        "gpu_launches": int(launches) + (cov_rec["gpu_launches"] if cov_rec else 0),
        "clocks": clocks,
        "host": {"cpus": os.cpu_count(), "placement_rank0": placement,
                 "topology": numa.topology_text() if world > 1 else ""},
    }
    if cov_rec:
        line["coverage"] = cov_rec

    if not args.no_cpu:
        import oracle

        oracle.build()
        threads = os.cpu_count() or 1
        m = min(args.mags, args.ref_mags)
        dt, b1, c_counts, c_gc = cpu_tnf_sample(oracle, hb, offsets, mag_off, m, threads)
        c1 = int(mag_off[m])
        ok = bool(np.array_equal(d_counts[:c1].cpu().numpy().view(np.uint32), c_counts)
                  and np.array_equal(d_gc[:c1].cpu().numpy(), c_gc, equal_nan=True))
        line["cpu_baseline"] = {"value": b1 / dt / 1e9, "unit": UNIT, "cores": threads, "kind": "port",
                                "fair_variant": fair_variant(oracle, hb, offsets, mag_off, m),
                                "sample": f"first {m} MAGs ({b1 / 1e6:.1f} Mbp) of this step's input; "
                                          f"get_tnf {threads} threads over contigs + serial 5-pass get_gc",
                                "parity_with_gpu": ok}
        if not ok:
            emit(line)
            raise SystemExit("PARITY FAILURE: GPU counts/GC differ from the CPU oracle")
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
