#!/usr/bin/env python
"""bench.py -- GCUPS / DO alignments per second of the batched affine alignment sweep
(BASELINE.json configs[1]: 100k pairs, lengths 500/2k/10k bp, Ukkonen band off and on).

One "step" = one pass of the hot path over the whole sweep: for every length, the cost-only
entry point on all pairs ("band off", batch twin of algn_CAML_cost_affine_3) and the banded
align + traceback + median entry point on all pairs ("band on", batch twin of
algn_CAML_align_affine_3), followed by the min-reduction of the candidate costs (NCCL
all-reduce when N > 1).  `value` counts full-matrix-equivalent cells (len_i-1)*(len_j-1) of every
alignment performed per second with inputs resident in HBM; `e2e` is the same
sweep through the host-buffer C ABI (pinned host inputs uploaded and all results read back
inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun ... bench.py --gpus N ...      (one rank per GPU, weak scaling: each rank owns
                                             its own --pairs pairs per length)
"""
import argparse
import ctypes
import json
from concurrent.futures import ThreadPoolExecutor
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np

METRIC = "GCUPS (full-matrix-equivalent cell updates/s) of batched affine DO alignment, band off + band on"
LENGTHS = (500, 2000, 10000)
REGIME = (1, 1, 3)          # R1: substitution 1, indel 1, gap opening 3
SEED = 0x504F5935
OPS_PER_CELL = {"gapfree": 9, "general": 16, "band": 30}   # scalar int32 ops per cost-only cell / per band cell with directions


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=100000, help="pairs per length per GPU")
    ap.add_argument("--lengths", default=",".join(str(x) for x in LENGTHS))
    ap.add_argument("--chunk-bases", type=int, default=600_000_000, help="max pool bytes per batch call")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled oracle replay of the timed pairs")
    ap.add_argument("--no-swap", action="store_true", help="skip the strong-scaled swap-evaluation sub-record")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0] / [2] / [4] sub-records (N = 1 only)")
    ap.add_argument("--swap-prunings", type=int, default=0, help="SPR prunings of the swap-evaluation workload (whole job); 0 = the full neighbourhood")
    ap.add_argument("--swap-chunk", type=int, default=128, help="prunings per candidate batch (bounds the node store of a rank)")
    ap.add_argument("--swap-check", type=int, default=24, help="medians / distances replayed on the CPU checker")
    return ap.parse_args()


def make_config(pairs, lengths, world, total_bytes=None):
    cfg = dict(workload="configs[1]: batched pairwise affine alignment sweep, %d pairs per length per GPU, lengths %s bp, "
                        "Ukkonen band off (cost-only) and on (align+traceback+median), 1 GPU per rank" % (pairs, list(lengths)),
               regime="subst %d indel %d gap_open %d" % REGIME,
               l2="inputs larger than L2 (%.1f GB of sequences per step)" % ((total_bytes or sum(2.0 * pairs * (L + 1) for L in lengths)) / 1e9),
               pairs_per_length=pairs, lengths=list(lengths),
               parallelism="pairs sharded over %d GPU(s), NCCL min-reduce of candidate costs" % world,
               streams="value: one context (stream) per GPU, the batches of a step back to back; e2e: two host threads with "
                       "one context each take alternate batches, so copies, host-side scheduling and kernel tails of one "
                       "batch overlap the kernels of the other (which is why e2e can exceed value)")
    return cfg


# ---------------------------------------------------------------------------------------------
def chunks_for(npairs, length, chunk_bases):
    """equal-sized batches of at most chunk_bases sequence bytes"""
    per = max(1, chunk_bases // (2 * (length + 8)))
    nchunks = (npairs + per - 1) // per
    out, p = [], 0
    for c in range(nchunks):
        n = (npairs - p + (nchunks - c) - 1) // (nchunks - c)
        out.append((p, n))
        p += n
    return out


def cpu_reference(lengths, threads, budget_core_s=24.0, seed=SEED):
    """Times the reference's own CPU implementation (oracle/_ref, unmodified src/algn.c) -- or the
    oracle port if the compiled reference is absent -- on a bounded sample of the same workload."""
    from poy5_b200 import synth
    from oracle import cost_matrix_oracle as cmo
    full, _ = cmo.dna_matrices(*REGIME)
    kind, lib, cm = None, None, None
    try:
        from oracle import refbind
        if refbind.available(True):
            lib = refbind.RefLib(True)
            cm = lib.cm(full)
            kind = "reference"
    except Exception:
        lib = None
    if lib is None:
        from oracle.port import Port
        lib = Port()
        cm = lib.cm(full)
        kind = "port"
    # Bounded sample with the SAME proportions as the GPU workload (equal pair counts per length, so the longest
    # length dominates the cells exactly as it does there): n pairs per length, n a multiple of the thread count,
    # sized from the probe rate of SURVEY.md section 6 (~0.17 GCUPS/core).
    per_pair_core_s = 2.0 * sum(float(L) * L for L in lengths) / 0.17e9
    n = max(1, int(round(budget_core_s / (threads * per_pair_core_s)))) * threads
    cells_tot, secs_tot, aln_tot, sample = 0, 0.0, 0, []
    for L in lengths:
        data, off = synth.pair_pool(seed + L, 0, n, L)
        lens = np.diff(off).astype(np.int32)
        la, lb = lens[0::2], lens[1::2]
        oa, ob = off[0:-1:2], off[1::2]
        swap = la > lb
        oi = np.where(swap, ob, oa); oj = np.where(swap, oa, ob)
        li = np.where(swap, lb, la).astype(np.int32); lj = np.where(swap, la, lb).astype(np.int32)
        cells = int(((la - 1).astype(np.int64) * (lb - 1)).sum())
        for mode in (0, 1):
            t, _ = lib.batch_affine(cm, mode, data, oi, li, oj, lj, swap.astype(np.uint8), threads)
            cells_tot += cells; secs_tot += t; aln_tot += n
        sample.append("%d pairs @%d" % (n, L))
    return dict(value=cells_tot / secs_tot / 1e9, unit="GCUPS", cores=threads, kind=kind,
                sample="band off + band on, " + ", ".join(sample) + " (same generator/seed as the GPU workload)",
                alignments_per_s=aln_tot / secs_tot, seconds=secs_tot)


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ---------------------------------------------------------------------------------------------
def emit(line):
    """the ONE JSON line goes to the real stdout; everything else that prints there (NCCL's version banner, ...) was
    redirected to stderr at start-up"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    lengths = tuple(int(x) for x in args.lengths.split(","))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = []
        for w in range(args.warmup + args.steps):
            r = cpu_reference(lengths, threads, budget_core_s=max(4.0, 1.5 * threads), seed=SEED + 7919 * w)
            if w >= args.warmup:
                steps.append(r)
        val = float(np.mean([r["value"] for r in steps])) if steps else 0.0
        r = steps[-1]
        line = dict(metric=METRIC, value=val, unit="GCUPS", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=1e3 * float(np.mean([x["seconds"] for x in steps])), higher_is_better=True,
                    scaling="weak", vs_baseline=None, dtype="int32", data="synthetic", impl="reference",
                    config=dict(make_config(args.pairs, lengths, max(1, args.gpus)),
                                reference_arm="POY5 algn.c on the host cores; each step is a bounded sample of this workload: " + r["sample"]),
                    cpu_baseline=dict(value=val, unit="GCUPS", cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    alignments_per_s=r["alignments_per_s"],
                    e2e=dict(value=val, unit="GCUPS", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        emit(line)
        return 0

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line (NCCL prints its version there)
    import torch
    import torch.distributed as dist
    import poy5_b200 as pb
    from poy5_b200 import synth, sequence
    from poy5_b200.cost_matrix import Two_D
    from poy5_b200.api import _ptr

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(dev)          # a real (non-legacy) stream shared by torch and the library
    torch.cuda.set_stream(stream)
    ctx = pb.Context(local_rank, stream.cuda_stream)
    t2d = Two_D.of_transformations_and_gaps(REGIME[0], REGIME[1], REGIME[2])
    cm = pb.CostModel(ctx, t2d.full)

    peak_ops, peak_clock = ctx.microbench(0)   # IADD3-class issue rate, measured live
    dpx_ops, _ = ctx.microbench(2)

    # ---- build the workload: host (pinned) + device-resident copies --------------------------------
    work = []   # one entry per (length, chunk)
    total_cells = 0
    total_aln = 0
    h2d = d2h = 0
    for L in lengths:
        for (p0, n) in chunks_for(args.pairs, L, args.chunk_bases):
            cap = synth.pair_pool_capacity(n, L)
            pinned = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            data, off = synth.pair_pool(SEED + L, rank * args.pairs + p0, n, L, out=pinned.numpy(), nthreads=min(threads, 16))
            lens = np.diff(off)
            ia = np.arange(0, 2 * n, 2, dtype=np.int32); ib = ia + 1
            la, lb = lens[ia], lens[ib]
            swaped = (la > lb).astype(np.uint8)
            si = np.where(swaped == 1, ib, ia).astype(np.int32); sj = np.where(swaped == 1, ia, ib).astype(np.int32)
            caps = (la + lb + 2).astype(np.int64)
            out_off = np.zeros(n, np.int64); np.cumsum(caps[:-1], out=out_off[1:])
            out_total = int(caps.sum())
            cells = int(((la - 1) * (lb - 1)).sum())
            w = dict(L=L, n=n, data=data, off=off, ia=ia, ib=ib, si=si, sj=sj, swaped=swaped, out_off=out_off,
                     out_total=out_total, cells=cells, pinned=pinned)
            # device-resident inputs/outputs for the `value` measurement
            w["d_data"] = torch.from_numpy(data).to(dev)
            w["d_off"] = torch.from_numpy(off).to(dev)
            w["d_ia"] = torch.from_numpy(ia).to(dev); w["d_ib"] = torch.from_numpy(ib).to(dev)
            w["d_sw"] = torch.from_numpy(swaped).to(dev)
            w["d_out_off"] = torch.from_numpy(out_off).to(dev)
            w["d_cost0"] = torch.empty(n, dtype=torch.int32, device=dev)
            w["d_cost1"] = torch.empty(n, dtype=torch.int32, device=dev)
            w["d_len"] = torch.empty(4 * n, dtype=torch.int32, device=dev)
            work.append(w)
            total_cells += 2 * cells
            total_aln += 2 * n
            h2d += data.nbytes + off.nbytes + 2 * (ia.nbytes + ib.nbytes) + swaped.nbytes + out_off.nbytes
            d2h += 2 * 4 * n + 16 * n + 4 * out_total
    max_out = max(w["out_total"] for w in work)
    d_outs = [torch.empty(max_out + 256, dtype=torch.uint8, device=dev) for _ in range(4)]
    # e2e: two host threads, each with its own context (stream, arenas) and pinned result buffers, take
    # alternate chunks, so that one chunk's copies and host-side scheduling overlap the other's kernels
    E2E_THREADS = 1 if args.no_e2e else 2
    lanes = []
    free_b, _total_b = torch.cuda.mem_get_info(dev)
    arena = int(min(48 << 30, 0.30 * free_b))      # direction-byte arena per context (two contexts share the HBM)
    for t in range(E2E_THREADS):
        c = ctx if t == 0 else pb.Context(local_rank)
        c.set_arena_limit(arena)
        lanes.append(dict(ctx=c, cm=cm if t == 0 else pb.CostModel(c, t2d.full),
                          h_outs=[torch.empty(max_out + 256, dtype=torch.uint8, pin_memory=True) for _ in range(4)]))
    best = torch.zeros(1, dtype=torch.int64, device=dev)

    seg_events = {}
    seg_band_cells = {}

    def local_min(costs):
        # candidate min-reduction inside the rank: (cost << 32 | index) packed int64, min over the batch
        packed = (costs.to(torch.int64) << 32) | torch.arange(costs.numel(), device=dev, dtype=torch.int64)
        return packed.min().reshape(1)

    def step_resident(record=False):
        best.fill_(torch.iinfo(torch.int64).max)
        for wi, w in enumerate(work):
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            pool = sequence.DevicePool(ctx, w["d_data"].data_ptr(), w["d_off"].data_ptr(), w["off"])
            sequence.cost_2_dev(ctx, cm, pool, w["n"], w["d_ia"].data_ptr(), w["d_ib"].data_ptr(), w["d_cost0"].data_ptr())
            if record:
                e1.record(stream)
            bc0 = ctx.stats()["band_cells"] if record else 0
            sequence.align_affine_3_dev(ctx, cm, pool, w["si"], w["sj"], w["d_sw"].data_ptr(), w["d_out_off"].data_ptr(),
                                        w["d_cost1"].data_ptr(), d_outs[0].data_ptr(), d_outs[1].data_ptr(),
                                        d_outs[2].data_ptr(), d_outs[3].data_ptr(), w["d_len"].data_ptr())
            if record:
                seg_band_cells[wi] = ctx.stats()["band_cells"] - bc0
            best.copy_(torch.minimum(best, torch.minimum(local_min(w["d_cost0"]), local_min(w["d_cost1"]))))
            if record:
                e2.record(stream)
                seg_events[wi] = (e0, e1, e2)
            pool.close()
        if world > 1:      # one MIN all-reduce of the packed (cost, candidate) per round, not per batch
            dist.all_reduce(best, op=dist.ReduceOp.MIN)

    def e2e_lane(lane, items):
        c, m, h_outs = lane["ctx"], lane["cm"], lane["h_outs"]
        res = 1 << 60
        for w in items:
            pool = pb.Pool(c, data=w["data"], offsets=w["off"])
            cost0 = sequence.Align.cost_2(c, m, pool, w["ia"], w["ib"])
            n = w["n"]
            cost1 = np.empty(n, np.int32); out_len = np.empty(4 * n, np.int32)
            c.check(c.L.poy_batch_align_affine(c.h, m.h, pool.h, n, _ptr(w["si"]), _ptr(w["sj"]), _ptr(w["swaped"]),
                                               _ptr(w["out_off"]), _ptr(cost1),
                                               ctypes.c_void_p(h_outs[0].data_ptr()), ctypes.c_void_p(h_outs[1].data_ptr()),
                                               ctypes.c_void_p(h_outs[2].data_ptr()), ctypes.c_void_p(h_outs[3].data_ptr()),
                                               _ptr(out_len), None))
            res = min(res, int(cost0.min()), int(cost1.min()))
            pool.close()
        return res

    def step_e2e():
        # longest chunks first, dealt alternately (the ctypes calls release the GIL)
        order = sorted(work, key=lambda w: -w["cells"])
        with ThreadPoolExecutor(len(lanes)) as ex:
            futs = [ex.submit(e2e_lane, lanes[t], order[t::len(lanes)]) for t in range(len(lanes))]
            res = min(f.result() for f in futs)
        if world > 1:
            t = torch.tensor([res], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            res = int(t.item())
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident (value) measurement ---------------------------------------------------------------
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = ctx.launches
    stats0 = ctx.stats()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for k in range(args.steps):
        step_resident(record=(k == args.steps - 1))
    ev1.record(stream)
    barrier()
    launches = ctx.launches - launches0
    stats1 = ctx.stats()
    band_cells_step = (stats1["band_cells"] - stats0["band_cells"]) / max(1, args.steps)
    fills_step = {k: (stats1[k] - stats0[k]) / max(1, args.steps) for k in ("probe_fills", "full_fills", "repeated", "rounds")}
    ms = ev0.elapsed_time(ev1) / max(1, args.steps)
    clocks = sampler.stop() if sampler else {}
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())

    # per-segment breakdown (last timed step)
    breakdown = []
    kernel_ms = {}
    band_ms = {}
    for wi, w in enumerate(work):
        if wi in seg_events:
            e0, e1, e2 = seg_events[wi]
            t_off, t_on = e0.elapsed_time(e1), e1.elapsed_time(e2)
            band_ms.setdefault(w["L"], [0.0, 0])
            band_ms[w["L"]][0] += t_on; band_ms[w["L"]][1] += seg_band_cells.get(wi, 0)
            breakdown.append(dict(L=w["L"], pairs=w["n"], band_off_ms=round(t_off, 3), band_on_ms=round(t_on, 3),
                                  band_on_cells_computed=int(seg_band_cells.get(wi, 0)),
                                  band_off_gcups=round(w["cells"] / t_off / 1e6, 2), band_on_gcups=round(w["cells"] / t_on / 1e6, 2),
                                  band_off_aln_per_s=round(w["n"] / t_off * 1e3, 1), band_on_aln_per_s=round(w["n"] / t_on * 1e3, 1)))
            kernel_ms.setdefault(w["L"], [0.0, 0])
            kernel_ms[w["L"]][0] += t_off; kernel_ms[w["L"]][1] += w["cells"]

    # ---- e2e measurement ------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        barrier()
        te = (time.perf_counter() - t0) / max(1, args.steps)
        tt = torch.tensor([te], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
        e2e = dict(value=world * total_cells / te / 1e9, unit="GCUPS", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                   ms_per_step=1e3 * te, alignments_per_s=world * total_aln / te)

    # ---- interior-node-like segment (not part of `value`): every pair carries ambiguity / gap-bit symbols, i.e. takes
    # the 4-state kernels that a tree pass runs --------------------------------------------------
    interior = None
    if not args.no_swap:
        Li, ni = 2000, min(20000, args.pairs)
        data, off = synth.pair_pool(SEED + 77, rank * ni, ni, Li, subst=0.15, indel=0.01, decorated=1.0, nthreads=min(threads, 16))
        lens = np.diff(off)
        ia = np.arange(0, 2 * ni, 2, dtype=np.int32); ib = ia + 1
        la, lb = lens[ia], lens[ib]
        sw = (la > lb).astype(np.uint8)
        si = np.where(sw == 1, ib, ia).astype(np.int32); sj = np.where(sw == 1, ia, ib).astype(np.int32)
        caps = (la + lb + 2).astype(np.int64)
        oo = np.zeros(ni, np.int64); np.cumsum(caps[:-1], out=oo[1:])
        cells_i = int(((la - 1) * (lb - 1)).sum())
        dd, do_ = torch.from_numpy(data).to(dev), torch.from_numpy(off).to(dev)
        dia, dib, dsw, doo = (torch.from_numpy(x).to(dev) for x in (ia, ib, sw, oo))
        dc0 = torch.empty(ni, dtype=torch.int32, device=dev); dc1 = torch.empty(ni, dtype=torch.int32, device=dev)
        dl = torch.empty(4 * ni, dtype=torch.int32, device=dev)
        outs_i = d_outs if int(caps.sum()) + 256 <= d_outs[0].numel() else [torch.empty(int(caps.sum()) + 256, dtype=torch.uint8, device=dev) for _ in range(4)]
        tms_i = []
        for rep in range(2):          # first pass warms up
            pool = sequence.DevicePool(ctx, dd.data_ptr(), do_.data_ptr(), off)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            sequence.cost_2_dev(ctx, cm, pool, ni, dia.data_ptr(), dib.data_ptr(), dc0.data_ptr())
            e1.record(stream)
            bc0 = ctx.stats()["band_cells"]
            sequence.align_affine_3_dev(ctx, cm, pool, si, sj, dsw.data_ptr(), doo.data_ptr(), dc1.data_ptr(), outs_i[0].data_ptr(),
                                        outs_i[1].data_ptr(), outs_i[2].data_ptr(), outs_i[3].data_ptr(), dl.data_ptr())
            e2.record(stream)
            torch.cuda.synchronize()
            bci = ctx.stats()["band_cells"] - bc0
            tms_i = [e0.elapsed_time(e1), e1.elapsed_time(e2)]
            pool.close()
        interior = dict(L=Li, pairs=ni, decorated=1.0, subst=0.15, band_off_ms=round(tms_i[0], 3), band_on_ms=round(tms_i[1], 3),
                        band_off_gcups=round(cells_i / tms_i[0] / 1e6, 2), band_on_gcups=round(cells_i / tms_i[1] / 1e6, 2),
                        band_on_cells_computed=int(bci), band_on_gcells_computed_per_s=round(bci / tms_i[1] / 1e6, 2),
                        note="100% interior-node-like pairs (4-state kernels); not part of `value`")
        del dd, do_, dia, dib, dsw, doo, dc0, dc1, dl, outs_i

    # ---- parity of the timed results: a sample of the timed pairs against the CPU checker (outside the timed region) -----
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import cost_matrix_oracle as cmo
        full, _ = cmo.dna_matrices(*REGIME)
        chk, kind = None, "port"
        try:
            from oracle import refbind
            if refbind.available(True):
                chk = refbind.RefLib(True); kind = "reference"
        except Exception:
            chk = None
        if chk is None:
            from oracle.port import Port
            chk = Port()
        pc = chk.cm(full)
        checked = mism = 0
        per_len = {500: 24, 2000: 10, 10000: 3}
        seen = set()
        for w in work:
            if w["L"] in seen:
                continue
            seen.add(w["L"])
            k = min(w["n"], per_len.get(w["L"], 4))
            idx = np.linspace(0, w["n"] - 1, k).astype(int)
            c0 = w["d_cost0"].cpu().numpy(); c1 = w["d_cost1"].cpu().numpy()      # results of the LAST timed step
            # the four sequences of the sampled pairs: re-run of exactly these pairs through the host C ABI
            sub = pb.Pool(ctx, [w["data"][w["off"][s]:w["off"][s + 1]] for q in idx for s in (2 * q, 2 * q + 1)])
            r = sequence.Align.align_affine_3(ctx, cm, sub, np.arange(0, 2 * k, 2, dtype=np.int32), np.arange(1, 2 * k, 2, dtype=np.int32))
            for j, q in enumerate(idx):
                a = w["data"][w["off"][2 * q]:w["off"][2 * q + 1]]; b = w["data"][w["off"][2 * q + 1]:w["off"][2 * q + 2]]
                swp = int(len(a) > len(b))
                xi, xj = (b, a) if swp else (a, b)
                oc, om, ow, ori, orj = chk.align_affine(pc, xi, xj, swp)
                ra, rb = (orj, ori) if swp else (ori, orj)
                ok = (int(c0[q]) == int(chk.cost_affine(pc, a, b)) and int(c1[q]) == int(oc) and int(r["cost"][j]) == int(oc)
                      and np.array_equal(om, r["median"][j]) and np.array_equal(ow, r["medianwg"][j])
                      and np.array_equal(ra, r["res_a"][j]) and np.array_equal(rb, r["res_b"][j]))
                checked += 1; mism += int(not ok)
            sub.close()
        parity = dict(checked=checked, mismatches=mism, against=kind,
                      what="sampled pairs of every length: cost-only cost and banded cost of the last timed step, plus median / "
                           "median_wg / both aligned rows of a re-run of the same pairs, vs the CPU checker",
                      unpinned="the Cost_matrix table fill (OCaml, cannot run here) and the tree-level enumerator are restatements "
                               "pinned only by self-consistency tests")

    # ---- swap evaluation: one SPR neighbourhood, strong-scaled over the ranks (north star) ------------------------------
    swap = None
    if not args.no_swap:
        from poy5_b200 import swap_eval
        for w in work:
            for k in [k for k in w if k.startswith("d_")]:
                del w[k]
        del d_outs
        for ln in lanes:
            ln["h_outs"] = None
        for ln in lanes[1:]:            # the second e2e lane holds tens of GB of direction arena: not needed any more
            ln["cm"].close(); ln["ctx"].close()
        del lanes[1:]
        ctx.trim()
        torch.cuda.empty_cache()
        try:
            swap, sample = swap_eval.run(ctx, rank=rank, world=world, device=dev, prunings=args.swap_prunings, chunk=args.swap_chunk,
                                         check=args.swap_check, regime=REGIME)
            if swap is not None and sample is not None:      # rank 0: CPU-checker replay of the recorded sample (untimed)
                from tests.oracle_backend import replay_sample
                t5 = time.perf_counter()
                swap["parity"] = replay_sample(sample[0], sample[1], REGIME)
                swap["parity"]["replay_s"] = time.perf_counter() - t5
        except Exception as e:          # reported, never hidden
            swap = dict(error="%s: %s" % (type(e).__name__, e))

    # ---- the other BASELINE configurations as short, parity-sampled sub-records (single GPU only).  They run in a child
    # process with a time-out: a failure of a side record must never take the headline line with it. -------------------
    configs = None
    if world == 1 and not args.no_configs and not args.no_swap:
        import pickle
        from tests.oracle_backend import replay_sample, replay_triplets, replay_newkk
        configs = {}
        tmp = tempfile.NamedTemporaryFile(suffix=".pkl", delete=False); tmp.close()
        torch.cuda.empty_cache()
        try:
            subprocess.run([sys.executable, "-m", "poy5_b200.workloads", "--out", tmp.name, "--device", str(local_rank),
                            "--regime", ",".join(str(x) for x in REGIME)], cwd=ROOT, timeout=240, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL)
        except Exception as e:
            configs["note"] = "child process: %s" % type(e).__name__
        try:
            with open(tmp.name, "rb") as f:
                got = pickle.load(f)
        except Exception:
            got = {}
        for name, val in got.items():
            if isinstance(val, dict):
                configs[name] = val; continue
            rec_, sample_ = val
            try:
                if sample_ is not None:
                    rec_["parity"] = (replay_triplets(sample_, REGIME) if name == "configs[2]" else replay_newkk(sample_, REGIME) if name == "newkkonen"
                                      else replay_sample(sample_[0], sample_[1], REGIME))
            except Exception as e:
                rec_["parity"] = dict(error="%s: %s" % (type(e).__name__, e))
            configs[name] = rec_
        try:
            os.unlink(tmp.name)
        except OSError:
            pass

    if rank == 0:
        # dominant kernels on the longest length: the cost-only wavefront (k_cost_affine) and the banded fill (k_band2)
        Ltop = max(kernel_ms) if kernel_ms else lengths[-1]
        k_ms, k_cells = kernel_ms.get(Ltop, (ms, total_cells))
        n_top = max(1, sum(1 for w in work if w["L"] == Ltop))   # launches of the dominant kernel per step
        k_bytes = sum(int(w["data"].nbytes) + 4 * w["n"] for w in work if w["L"] == Ltop) / n_top   # per launch
        hbm_peak = 6550.7
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        achieved = k_cells * OPS_PER_CELL["gapfree"] / (k_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the same shape, from this round's committed ncu capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["k_cost_affine"]
            top = [w for w in work if w["L"] == Ltop]
            if tj["length"] == Ltop and all(w["n"] == tj["pairs_per_launch"] for w in top):
                traffic = tj["dram_bytes_per_launch"]; traffic_src = "profiles/r02_traffic.json (ncu --set full of this kernel at this shape)"
        except Exception:
            pass
        b_ms, b_cells = band_ms.get(Ltop, (0.0, 0))
        band_roof = None
        if b_ms > 0 and b_cells > 0:
            b_ach = b_cells * OPS_PER_CELL["band"] / (b_ms * 1e-3) / 1e12
            band_roof = dict(bound="int32", kernel="k_band2 family (fills of the threshold-doubling schedule) + k_traceback",
                             achieved=b_ach, peak=peak_ops / 1e12, unit="Tops/s", frac=b_ach / (peak_ops / 1e12),
                             cells_computed=int(b_cells), gcells_computed_per_s=b_cells / (b_ms * 1e-3) / 1e9,
                             note="band cells computed over ALL fills of the band-on calls on the L=%d segments (probe fills and "
                                  "fills with direction bytes alike) x %d algorithmic ops per traceback cell (SURVEY 8d) / the "
                                  "band-on time of those segments (CUDA events; includes stop rule, traceback and the host-driven "
                                  "round trips), vs the same live IADD3-class issue rate" % (Ltop, OPS_PER_CELL["band"]))
        roofline = dict(bound="int32", kernel="k_cost_affine", achieved=achieved, peak=peak_ops / 1e12, unit="Tops/s",
                        frac=achieved / (peak_ops / 1e12), traffic=traffic, traffic_source=traffic_src,
                        note="INT32 issue roofline: algorithmic scalar add/min per cell (%d, gap-free cost-only cell) x cells / "
                             "launch time, vs the IADD3-class issue rate measured live by poy_microbench_int "
                             "(DPX VIADDMNMX measured %.2f Tops/s). Launch time from CUDA events around the cost-only "
                             "calls on the L=%d segment (about 10%% of its pairs carry gap bits and run the 4-state path, 16 ops/cell, but are counted at 9)."
                             % (OPS_PER_CELL["gapfree"], dpx_ops / 1e12, Ltop),
                        sm_clock_mhz_microbench=peak_clock, band=band_roof,
                        hbm=dict(algorithmic_bytes_per_launch=int(k_bytes), achieved_gbs=k_bytes * n_top / (k_ms * 1e-3) / 1e9,
                                 peak_gbs=hbm_peak, frac=k_bytes * n_top / (k_ms * 1e-3) / 1e9 / hbm_peak,
                                 note="sequence bytes in + 4 B cost out per pair (SURVEY 8d): this path is integer-issue bound, not HBM bound"))
        cpu = None
        if not args.no_cpu and world == 1:      # the CPU baseline leg runs at N = 1 only
            try:
                cpu = cpu_reference(lengths, threads)
                cpu = dict(value=cpu["value"], unit=cpu["unit"], cores=cpu["cores"], kind=cpu["kind"], sample=cpu["sample"],
                           alignments_per_s=cpu["alignments_per_s"])
            except Exception as e:  # the oracle always exists; report rather than hide
                cpu = dict(value=None, unit="GCUPS", cores=threads, kind="unavailable", sample=str(e))
        line = dict(metric=METRIC, value=world * total_cells / (ms * 1e-3) / 1e9, unit="GCUPS", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="int32", data="synthetic",
                    config=make_config(args.pairs, lengths, world, sum(w["data"].nbytes for w in work)),
                    alignments_per_s=world * total_aln / (ms * 1e-3), breakdown=breakdown,
                    cells_computed=dict(band_on_per_step_per_gpu=int(band_cells_step), band_off_per_step_per_gpu=int(total_cells // 2),
                                        fills_per_step={k: int(v) for k, v in fills_step.items()},
                                        note="band on: cells inside the Ukkonen bands summed over every fill of the threshold-doubling "
                                             "schedule (poy_ctx_stats); band off: the full matrices"),
                    breakdown_interior=interior, clocks=clocks, e2e=e2e,
                    gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, parity=parity, swap_eval=swap,
                    other_configs=configs)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
