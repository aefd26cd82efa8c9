#!/usr/bin/env python
"""bench.py -- the hot path on BASELINE.json's config 3 (`zipf1g`): a synthetic 1 GiB
Zipf-skewed order-0 byte stream per GPU, static coder, 64 KiB blocks.

One "step" = encode the stream into a B2RC container, then decode it back.
    metric  roundtrip_GBps = uncompressed bytes / (t_encode + t_decode), whole job
    value   inputs already resident in HBM (CUDA events, max over ranks)
    e2e     the same through the host-pointer C ABI (b2rc_encode / b2rc_decode) with
            pinned HOST buffers: H2D and D2H copies inside the timed region
Weak scaling: every rank codes its own 1 GiB shard (blocks shard by contiguous range,
cpprcoder_b200/dist.py); the only collective is the all-gather of payload sizes.

    python bench.py [--gpus N --steps K --warmup W]            our arm
    python bench.py --workload NAME                            other streams / coders of the same path, the rANS
                                                               sibling, the block-sort transform (WORKLOADS below)
    python bench.py --impl reference [...]                      the reference's CPU coder
under torchrun for N > 1 (one rank per GPU, NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
# The host-pointer pipeline keeps one stream per chunk busy (up to 16 + 2).  With the default
# of 8 hardware queues, streams alias and one chunk's copy waits behind another chunk's
# kernel; must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (generator, bytes per GPU, mode, block)
    "zipf1g-static-64k": ("zipf", 1 << 30, 0, 65536),
    "mixed-adaptive-64k": ("mixed", 1 << 30, 1, 65536),
    "kennedy-static-64k": ("kennedy", 1 << 30, 0, 65536),
    # the sibling rANS coder of the reference (cppans.h), eight interleaved states
    "zipf1g-rans-word-64k": ("zipf", 1 << 30, 3, 65536),
    "mixed-rans-word-64k": ("mixed", 1 << 30, 3, 65536),
    # the reference's block-sort transform (blksort.h): fixed 32 KiB blocks, step = forward + inverse
    "zipf1g-blksort": ("zipf", 1 << 30, 4, 32768),
    "mixed-blksort": ("mixed", 1 << 30, 4, 32768),
}
BLKSORT = 4
KERNEL_NAMES = {0: ("k_enc_static", "k_dec_static_seg"), 1: ("k_enc_adaptive", "k_dec_adaptive"),
                2: ("k_ans_enc_byte", "k_ans_dec_byte"), 3: ("k_ans_enc_word", "k_ans_dec_word"),
                4: ("k_blk_fwd", "k_blk_inv")}
METRIC = "roundtrip_GBps"
UNIT = "GB/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="zipf1g-static-64k", choices=sorted(WORKLOADS))
    ap.add_argument("--bytes", type=int, default=0, help="override bytes per GPU (testing only)")
    ap.add_argument("--cpu-sample", type=int, default=256 << 20, help="bytes of the stream the CPU baseline codes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def config_of(args, world):
    gen, nbytes, mode, block = WORKLOADS[args.workload]
    if args.bytes:
        nbytes = args.bytes
    return {
        "workload": args.workload,
        "generator": f"cpprcoder_b200.synth.{gen}",
        "coder": {0: "static", 1: "adaptive", 2: "rans-byte", 3: "rans-word", 4: "blksort (transform, no coder)"}[mode],
        "block_size": block,
        "bytes_per_gpu": nbytes,
        "global_bytes": nbytes * world,
        "parallelism": f"blocks sharded by contiguous range over {world} GPU(s)",
        "l2": "inputs (1 GiB per GPU) are larger than the 126 MB L2; no explicit flush",
        "restart_points": ("every 8192 symbols (static coder: 84 B per 64 KiB block behind the payloads, counted in "
                           "compressed_ratio; the payloads are the reference's)") if mode == 0 else "none",
    }, gen, nbytes, mode, block


# ------------------------------------------------------------------ clocks --
class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, device_index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            p = torch.cuda.get_device_properties(device_index)
            bus = "%08x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.nv = pynvml
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self.ok:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML sampling unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------- CPU arms --
def cpu_coder():
    """The reference's own CPU coder (oracle/_ref) when it was built, else the oracle port."""
    sys.path.insert(0, str(ROOT / "tests"))
    from _oracle import Oracle, Ref, offsets_of
    if Ref.available():
        ref = Ref.get()
        return ref, "reference", ref.hardware_threads(), offsets_of
    return Oracle.get(), "port", os.cpu_count() or 1, offsets_of


def cpu_roundtrip(coder, offsets_of, data, mode, block, threads):
    if mode == BLKSORT:  # blksort::BlkSort::encode / ::decode, blocks spread over `threads` BlkSort objects
        from _oracle import BlkSort
        bs = BlkSort(coder)
        t0 = time.perf_counter()
        coded = bs.encode(data, threads=threads)
        t1 = time.perf_counter()
        back = bs.decode(coded, threads=threads)
        t2 = time.perf_counter()
        assert back.tobytes() == data.tobytes(), "CPU baseline failed to round-trip"
        return t1 - t0, t2 - t1, int(coded.size)
    t0 = time.perf_counter()
    pays = coder.encode_blocks(mode, data, block, threads=threads)
    t1 = time.perf_counter()
    stream = np.frombuffer(b"".join(pays), dtype=np.uint8)
    off = offsets_of(pays)
    t2 = time.perf_counter()
    back = coder.decode_blocks(mode, stream, off, block, data.size, threads=threads)
    t3 = time.perf_counter()
    assert back.tobytes() == data.tobytes(), "CPU baseline failed to round-trip"
    return t1 - t0, t3 - t2, int(stream.size)


def cpu_baseline(args, gen, nbytes, mode, block):
    from cpprcoder_b200 import synth
    coder, kind, cores, offsets_of = cpu_coder()
    sample = min(args.cpu_sample, nbytes)
    sample -= sample % block
    data = synth.GENERATORS[gen](sample)
    te, td, c = cpu_roundtrip(coder, offsets_of, data, mode, block, cores)
    small = data[:min(sample, 32 << 20)]
    te1, td1, _ = cpu_roundtrip(coder, offsets_of, small, mode, block, 1)
    return {"value": sample / (te + td) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {sample >> 20} MiB of the same stream, block-parallel over {cores} host threads",
            "encode_GBps": sample / te / 1e9, "decode_GBps": sample / td / 1e9,
            "single_thread": {"value": small.size / (te1 + td1) / 1e9, "encode_GBps": small.size / te1 / 1e9,
                              "decode_GBps": small.size / td1 / 1e9, "sample": f"first {small.size >> 20} MiB"}}


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path, all host threads,
    each step a bounded sample of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from cpprcoder_b200 import synth
    cfg, gen, nbytes, mode, block = config_of(args, world)
    coder, kind, cores, offsets_of = cpu_coder()
    sample = min(128 << 20, nbytes)
    sample -= sample % block
    data = synth.GENERATORS[gen](sample)
    for _ in range(args.warmup):
        cpu_roundtrip(coder, offsets_of, data[:min(sample, 16 << 20)], mode, block, cores)
    te = td = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        a, b, _ = cpu_roundtrip(coder, offsets_of, data, mode, block, cores)
        te += a
        td += b
    wall = time.perf_counter() - t0
    value = sample * args.steps / (te + td) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * (te + td) / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic", "impl": "reference",
            "config": cfg, "gpu_launches": 0,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"each step codes the first {sample >> 20} MiB of the stream on {cores} host threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "encode_GBps": sample * args.steps / te / 1e9, "decode_GBps": sample * args.steps / td / 1e9,
            "wall_s": wall}
    print(json.dumps(line), flush=True)


class _StdoutToStderr:
    """NCCL prints its version banner with a C-level printf to stdout when the communicator comes up.
    stdout carries ONE JSON line, so file descriptor 1 points at stderr while that can happen."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


# ------------------------------------------------------------------ our arm --
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cpprcoder_b200 import api, container, synth
    from cpprcoder_b200 import dist as rcdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("N > 1 runs under torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path to time")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()  # brings the communicator up (and its banner out) here

    cfg, gen, nbytes, mode, block = config_of(args, world)
    n_total = nbytes * world
    lo, hi, blk_lo, blk_hi = rcdist.shard_of(n_total, block, rank, world)
    data = synth.GENERATORS[gen](hi - lo, start=lo)  # this rank's bytes of the global stream
    n = data.size
    ctx = api.Context(local)
    src = torch.from_numpy(data).to(dev)
    bound = api.blk_encode_bound(n) if mode == BLKSORT else api.bound(mode, n, block)
    enc = torch.empty(bound, dtype=torch.uint8, device=dev)
    dec = torch.empty(n, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    state = {}

    def step_device(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timers is not None else None
        if ev:
            ev[0].record()
        if mode == BLKSORT:  # sizes are a function of n: the shards need no exchange at all
            state["used"] = ctx.blk_encode_device(src, enc).numel()
            if ev:
                ev[1].record()
            ctx.blk_decode_device(enc, state["used"], dec)
        elif world == 1:
            _, used = ctx.encode_device(mode, src, enc, block)
            state["used"] = used
            if ev:
                ev[1].record()
            ctx.decode_device(enc, used, dec)
        else:
            shard = rcdist.encode_shard(ctx, mode, src, n_total, block)  # K1 K2 K4 + all-gather of sizes (NCCL)
            state["shard"] = shard
            state["used"] = shard.payload_bytes
            if ev:
                ev[1].record()
            rcdist.decode_shard(ctx, shard, dec)
        if ev:
            ev[2].record()
            timers.append(ev)

    # ---- warm-up, then a correctness gate (a number for wrong bytes is worthless)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    assert torch.equal(dec, src), "round trip failed: refusing to report throughput"
    comp_bytes = state["used"]

    # ---- timed: device resident
    ctx.profile(True)
    ksum = {}
    timers = []
    sampler = ClockSampler(local)
    launches0 = ctx.launches
    barrier()
    sampler.start()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for _ in range(args.steps):
        step_device(timers)
        for k, v in ctx.kernel_ms().items():
            ksum.setdefault(k, []).append(v)
    t_end.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    ctx.profile(False)
    total_ms = t_start.elapsed_time(t_end)
    enc_ms = sum(e[0].elapsed_time(e[1]) for e in timers)
    dec_ms = sum(e[1].elapsed_time(e[2]) for e in timers)
    times = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms = (float(x) for x in times.tolist())
    ms_per_step = total_ms / args.steps
    value = n_total * args.steps / (total_ms / 1e3) / 1e9  # all ranks' bytes / max-over-ranks time

    # ---- timed: end to end through the host-pointer C ABI, pinned host buffers
    e2e = None
    if not args.no_e2e:
        h_src = torch.from_numpy(data).pin_memory()
        h_enc = torch.empty(bound, dtype=torch.uint8).pin_memory()
        h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
        a_src, a_enc, a_dec = h_src.numpy(), h_enc.numpy(), h_dec.numpy()

        def step_host():
            if mode == BLKSORT:
                out = ctx.blk_encode(a_src, dst=a_enc)                 # H2D n, kernels, D2H coded
                ctx.blk_decode(out, dst=a_dec)                         # H2D coded, kernels, D2H n
                return out.size
            out = ctx.encode(mode, a_src, block, dst=a_enc)           # H2D n, kernels, D2H container
            if world > 1:  # the stitched index needs every rank's sizes: the same small collective
                info = container.parse(out)
                sizes = torch.from_numpy(np.diff(info.offsets.astype(np.int64)).astype(np.int32)).to(dev)
                rcdist.allgather_sizes(sizes, n_total, block)
            ctx.decode(out, dst=a_dec)                                 # H2D container, kernels, D2H n
            return out.size

        for _ in range(2):
            used_host = step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            used_host = step_host()
        barrier()
        wall = time.perf_counter() - t0
        assert bytes(a_dec[:4096]) == bytes(data[:4096]) and bytes(a_dec[-4096:]) == bytes(data[-4096:])
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
        e2e = {"value": n_total * args.steps / wall / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n + used_host),
               "d2h_bytes_per_step": int(used_host + n), "ms_per_step": 1e3 * wall / args.steps,
               "api": ("b2rc_blk_encode + b2rc_blk_decode" if mode == BLKSORT else "b2rc_encode + b2rc_decode") +
                      " (host pointers, pinned)"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        kavg = {k: float(np.mean(v)) for k, v in ksum.items()}
        algo = {"histogram": n + (512 if mode == 0 else 1032) * (blk_hi - blk_lo), "encode": n + comp_bytes,
                "scan": 12 * (blk_hi - blk_lo),
                "compact": 2 * comp_bytes, "decode": comp_bytes + n,
                "blk_forward": n + comp_bytes, "blk_inverse": comp_bytes + n}
        kernels = {k: {"ms": ms, "algorithmic_bytes": algo[k], "GBps": algo[k] / ms / 1e6, "hbm_frac": algo[k] / ms / 1e6 / peak}
                   for k, ms in kavg.items() if k in algo and ms > 0}
        fwd_k, inv_k = ("blk_forward", "blk_inverse") if mode == BLKSORT else ("encode", "decode")
        dom = max((k for k in kernels if k in (fwd_k, inv_k)), key=lambda k: kernels[k]["ms"]) if kernels else None
        traffic = None
        try:  # per-launch DRAM bytes from the committed ncu capture, when there is one for this kernel
            tj = json.loads((ROOT / "profiles" / "dram_traffic.json").read_text())
            traffic = tj.get(f"{args.workload}:{dom}")
        except Exception:
            pass
        roofline = None
        if dom:
            roofline = {"kernel": KERNEL_NAMES[mode][0 if dom == fwd_k else 1],
                        "bound": "hbm", "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                        "frac": kernels[dom]["hbm_frac"], "traffic": traffic, "peak_source": peak_src,
                        "note": ("shared-memory bound sort kernel (one 32 KiB block per CTA, every pass a gather and a "
                                 "scatter through shared memory); HBM fraction shown for context") if mode == BLKSORT else
                                ("integer-pipe / latency bound coder kernel (serial chains, ~100 instructions per "
                                 "symbol); HBM fraction shown for context, issue utilisation is in profiles/")}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "encode_GBps": n_total * args.steps / (enc_ms / 1e3) / 1e9,
                "decode_GBps": n_total * args.steps / (dec_ms / 1e3) / 1e9,
                "compressed_ratio": comp_bytes / n, "kernels": kernels}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, gen, nbytes, mode, block)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
