#!/usr/bin/env python
"""bench.py -- the hot path on BASELINE.json's config 3 (`zipf1g`): a synthetic 1 GiB
Zipf-skewed order-0 byte stream per GPU, static coder, 64 KiB blocks.

One "step" = encode the stream into a B2RC container, then decode it back.
    metric  roundtrip_GBps = uncompressed bytes / (t_encode + t_decode), whole job
    value   inputs already resident in HBM (CUDA events, max over ranks); WEAK scaling: every rank
            codes its own 1 GiB shard of an N GiB stream (blocks shard by contiguous range,
            cpprcoder_b200/dist.py; the only collective is the all-gather of payload sizes)
    e2e     the same through the host-pointer C ABI (b2rc_encode / b2rc_decode) with
            pinned HOST buffers: H2D and D2H copies inside the timed region
    strong  BASELINE config 3 as written: ONE 1 GiB stream cut over the N ranks
    parity  before any number is printed, payloads are compared byte for byte with the
            reference's own encoder run on the same blocks (every rank, its first blocks)
    extra   the adaptive coder on the mixed stream (config 4's per-GPU share), the static coder at
            1 MiB blocks (config 5's far end), decode of a container without restart points

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
    "zipf1g-static-1m": ("zipf", 1 << 30, 0, 1 << 20),
    # the sibling rANS coder of the reference (cppans.h), eight interleaved states
    "zipf1g-rans-word-64k": ("zipf", 1 << 30, 3, 65536),
    "mixed-rans-word-64k": ("mixed", 1 << 30, 3, 65536),
    # the reference's block-sort transform (blksort.h): fixed 32 KiB blocks, step = forward + inverse
    "zipf1g-blksort": ("zipf", 1 << 30, 4, 32768),
    "mixed-blksort": ("mixed", 1 << 30, 4, 32768),
}
BLKSORT = 4
CODERS = {0: "static", 1: "adaptive", 2: "rans-byte", 3: "rans-word", 4: "blksort (transform, no coder)"}
KERNEL_NAMES = {0: ("k_enc_seg", "k_dec_static_seg"), 1: ("k_enc_adaptive", "k_dec_adaptive"),
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
    ap.add_argument("--parity-blocks", type=int, default=1024,
                    help="blocks per rank whose payloads are compared with the reference's (N = 1: the CPU sample's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra legs (adaptive / 1 MiB blocks / no restart)")
    return ap.parse_args()


def config_of(args, world):
    gen, nbytes, mode, block = WORKLOADS[args.workload]
    if args.bytes:
        nbytes = args.bytes
    return {
        "workload": args.workload,
        "generator": f"cpprcoder_b200.synth.{gen}",
        "coder": CODERS[mode],
        "block_size": block,
        "bytes_per_gpu": nbytes,
        "global_bytes": nbytes * world,
        "parallelism": f"blocks sharded by contiguous range over {world} GPU(s)",
        "l2": "inputs (1 GiB per GPU) are larger than the 126 MB L2; no explicit flush",
        "restart_points": ("every 4096 symbols for a stream that fills the GPU at that spacing, as 1 GiB per GPU does "
                           "(static coder: 180 B per 64 KiB block behind the payloads, counted in compressed_ratio; the "
                           "payloads are the reference's); closer together for fewer blocks (b2rc_restart_for)") if mode == 0
                          else ("adaptive coder: every 21888 symbols, 524 B a point with the model's counts, counted in "
                                "compressed_ratio" if mode == 1 else "none"),
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


def bind_to_gpu_numa(device_index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any host buffer is
    allocated: pinned pages are then node-local (first touch) and the N ranks of a box do not all
    pull their PCIe traffic through one socket's memory.  Returns a note for the JSON line."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        node = int(Path(f"/sys/bus/pci/devices/{bus}/numa_node").read_text())
        if node < 0:
            return "numa_node -1 (single node or not exposed): not bound"
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return f"numa node {node}: none of its CPUs is allowed to this process"
        os.sched_setaffinity(0, cpus)
        return f"bound to numa node {node} ({len(cpus)} CPUs)"
    except Exception as e:
        return f"not bound ({type(e).__name__})"


# --------------------------------------------------------------- CPU arms --
def cpu_coder():
    """The reference's own CPU coder (oracle/_ref) when it was built, else the oracle port."""
    sys.path.insert(0, str(ROOT / "tests"))
    from _oracle import Oracle, Ref, offsets_of
    if Ref.available():
        ref = Ref.get()
        return ref, "reference", ref.hardware_threads(), offsets_of
    return Oracle.get(), "port", os.cpu_count() or 1, offsets_of


def cpu_roundtrip(coder, offsets_of, data, mode, block, threads, keep=None):
    if mode == BLKSORT:  # blksort::BlkSort::encode / ::decode, blocks spread over `threads` BlkSort objects
        from _oracle import BlkSort
        bs = BlkSort(coder)
        t0 = time.perf_counter()
        coded = bs.encode(data, threads=threads)
        t1 = time.perf_counter()
        back = bs.decode(coded, threads=threads)
        t2 = time.perf_counter()
        assert back.tobytes() == data.tobytes(), "CPU baseline failed to round-trip"
        if keep is not None:
            keep["coded"] = coded
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
    if keep is not None:
        keep["stream"], keep["offsets"] = stream, off
    return t1 - t0, t3 - t2, int(stream.size)


def cpu_baseline(coder_tuple, data, mode, block, keep):
    """The reference on this box's host cores over `data` (block-parallel, all threads) and, on a
    shorter prefix, one thread.  `keep` receives the payloads: the parity gate compares them."""
    coder, kind, cores, offsets_of = coder_tuple
    te, td, c = cpu_roundtrip(coder, offsets_of, data, mode, block, cores, keep)
    small = data[:min(data.size, 32 << 20)]
    te1, td1, _ = cpu_roundtrip(coder, offsets_of, small, mode, block, 1)
    return {"value": data.size / (te + td) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {data.size >> 20} MiB of the same stream, block-parallel over {cores} host threads",
            "encode_GBps": data.size / te / 1e9, "decode_GBps": data.size / td / 1e9,
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
    cfg["reference_sample"] = (f"every step codes the first {sample >> 20} MiB of the stream (bounded: the CPU coder "
                               f"needs seconds per GiB); ms_per_step is per {sample >> 20} MiB, not per bytes_per_gpu")
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


def make_ctx(api, device, **env):
    """A context created under extra B2RC_* environment (the library reads it at creation)."""
    saved = {k: os.environ.get(k) for k in env}
    try:
        for k, v in env.items():
            os.environ[k] = str(v)
        return api.Context(device)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


# ------------------------------------------------------------------ our arm --
class Leg:
    """One stream resident on this rank's GPU and the device-timed round trip over it."""

    def __init__(self, env, gen, n_total, mode, block, ctx=None):
        import torch
        from cpprcoder_b200 import api, synth
        from cpprcoder_b200 import dist as rcdist
        self.env, self.mode, self.block, self.n_total = env, mode, block, n_total
        self.ctx = ctx or env["ctx"]
        self.lo, self.hi, self.blk_lo, self.blk_hi = rcdist.shard_of(n_total, block, env["rank"], env["world"])
        self.data = synth.GENERATORS[gen](self.hi - self.lo, start=self.lo)  # this rank's bytes of the global stream
        self.n = self.data.size
        self.src = torch.from_numpy(self.data).to(env["dev"])
        self.bound = api.blk_encode_bound(self.n) if mode == BLKSORT else api.bound(mode, self.n, block)
        self.enc = torch.empty(max(self.bound, 16), dtype=torch.uint8, device=env["dev"])
        self.dec = torch.empty(max(self.n, 16), dtype=torch.uint8, device=env["dev"])
        self.used = 0
        self.shard = None

    def step(self, ev=None, dec_ctx=None):
        from cpprcoder_b200 import dist as rcdist
        ctx, mode = self.ctx, self.mode
        if ev:
            ev[0].record()
        if mode == BLKSORT:  # sizes are a function of n: the shards need no exchange at all
            self.used = ctx.blk_encode_device(self.src, self.enc).numel()
            if ev:
                ev[1].record()
            ctx.blk_decode_device(self.enc, self.used, self.dec)
        elif self.env["world"] == 1:
            _, self.used = ctx.encode_device(mode, self.src, self.enc, self.block)
            if ev:
                ev[1].record()
            (dec_ctx or ctx).decode_device(self.enc, self.used, self.dec)
        else:
            # one b2rc_encode_device call per rank + the all-gather of payload sizes (NCCL).  A rank decodes its
            # own blocks from its own index, so the decode is launched first and the collective behind it
            # (B2RC_BENCH_GATHER=overlap: the collective first, running beside the decode -- measured slower:
            # its kernel sits on SMs the decoder's single wave needs)
            overlap = os.environ.get("B2RC_BENCH_GATHER", "after") == "overlap"
            self.shard = rcdist.encode_shard(ctx, mode, self.src, self.n_total, self.block, dst=self.enc,
                                             defer_gather=not overlap)
            self.used = self.shard.used
            if ev:
                ev[1].record()
            rcdist.decode_shard(dec_ctx or ctx, self.shard, self.dec)
            self.shard.offsets  # the global index is part of the step: the collective has to have finished
        if ev:
            ev[2].record()

    def run(self, warmup, steps, profile=False):
        """W untimed + K timed steps; times are the max over ranks."""
        import torch
        import torch.distributed as dist
        env, ctx = self.env, self.ctx
        for _ in range(warmup):
            self.step()
        env["barrier"]()
        assert torch.equal(self.dec[:self.n], self.src), "round trip failed: refusing to report throughput"
        ksum, timers = {}, []
        if profile:
            ctx.profile(True)
        launches0 = ctx.launches
        env["barrier"]()
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record()
        for _ in range(steps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            self.step(ev)
            timers.append(ev)
            if profile:
                for k, v in ctx.kernel_ms().items():
                    ksum.setdefault(k, []).append(v)
        t_end.record()
        env["barrier"]()
        launches = ctx.launches - launches0
        if profile:
            ctx.profile(False)
        total_ms = t_start.elapsed_time(t_end)
        enc_ms = sum(e[0].elapsed_time(e[1]) for e in timers)
        dec_ms = sum(e[1].elapsed_time(e[2]) for e in timers)
        times = torch.tensor([total_ms, enc_ms, dec_ms], dtype=torch.float64, device=env["dev"])
        if env["world"] > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        total_ms, enc_ms, dec_ms = (float(x) for x in times.tolist())
        gb = self.n_total * steps / 1e9
        return {"value": gb / (total_ms / 1e3), "ms_per_step": total_ms / steps, "encode_GBps": gb / (enc_ms / 1e3),
                "decode_GBps": gb / (dec_ms / 1e3), "compressed_ratio": self.used / max(self.n, 1),
                "kernel_ms": {k: float(np.mean(v)) for k, v in ksum.items()}, "launches": int(launches)}

    def parity(self, coder_tuple, nblocks, ref=None):
        """Compares the first `nblocks` payloads of this rank's container, byte for byte, with the
        reference's own encoder run on the same blocks (test/main.cpp:295-299 compares every byte
        too), and the index with their sizes.  Returns the number of blocks that were compared."""
        from cpprcoder_b200 import container
        coder, kind, cores, offsets_of = coder_tuple
        have = self.blk_hi - self.blk_lo
        nb = min(nblocks, have)
        if nb == 0:
            return 0
        if self.mode == BLKSORT:
            from _oracle import BlkSort
            nbytes = min(nb * 32768, self.n)
            want = ref["coded"] if ref and "coded" in ref else BlkSort(coder).encode(self.data[:nbytes], threads=cores)
            nb = nbytes // 32768
            got = self.enc[:nb * 32770].cpu().numpy()
            assert got.tobytes() == want[:nb * 32770].tobytes(), "block sort output differs from the reference's"
            return nb
        idx = container.HEADER + 8 * (have + 1)
        head = self.enc[:idx].cpu().numpy()
        offs = np.frombuffer(head[container.HEADER:].tobytes(), dtype=np.uint64)
        if ref and "stream" in ref:
            stream, want_off = ref["stream"], ref["offsets"]
            nb = min(nb, len(want_off) - 1)
        else:
            sample = self.data[:min(nb * self.block, self.n)]
            pays = coder.encode_blocks(self.mode, sample, self.block, threads=max(1, cores // self.env["local_world"]))
            stream, want_off = np.frombuffer(b"".join(pays), dtype=np.uint8), offsets_of(pays)
        assert offs[0] == 0 and np.array_equal(offs[:nb + 1], np.asarray(want_off[:nb + 1], dtype=np.uint64)), \
            "payload sizes differ from the reference's"
        end = int(offs[nb])
        got = self.enc[idx:idx + end].cpu().numpy()
        if got.tobytes() != stream[:end].tobytes():
            bad = next(b for b in range(nb) if got[int(offs[b]):int(offs[b + 1])].tobytes()
                       != stream[int(offs[b]):int(offs[b + 1])].tobytes())
            raise AssertionError(f"rank {self.env['rank']}: payload of block {self.blk_lo + bad} differs from the "
                                 f"reference's: refusing to report throughput")
        if self.shard is not None:  # N > 1: the replicated global index agrees with this rank's own
            g = self.shard.offsets[self.blk_lo:self.blk_lo + nb + 1].cpu().numpy().astype(np.int64)
            assert np.array_equal(g - g[0], offs[:nb + 1].astype(np.int64)), "gathered index differs from the local one"
        return nb


def copy_ceiling(env, n_bytes, c_bytes, h_a, h_b, steps=3):
    """What the box allows the e2e step, kernels aside: H2D of the input while the container goes
    D2H (encode), then H2D of the container while the output goes D2H (decode), plain
    cudaMemcpyAsync on two streams from the SAME pinned buffers.  Max over ranks."""
    import torch
    import torch.distributed as dist
    dev = env["dev"]
    d_a = torch.empty(n_bytes, dtype=torch.uint8, device=dev)
    d_b = torch.empty(c_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def once():
        for up, down in (((d_a, h_a[:n_bytes]), (h_b[:c_bytes], d_b)), ((d_b, h_b[:c_bytes]), (h_a[:n_bytes], d_a))):
            with torch.cuda.stream(s1):
                up[0].copy_(up[1], non_blocking=True)
            with torch.cuda.stream(s2):
                down[0].copy_(down[1], non_blocking=True)
            s1.synchronize()
            s2.synchronize()

    once()
    env["barrier"]()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    env["barrier"]()
    wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if env["world"] > 1:
        dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    return float(wall.item()) / steps


def run_ours(args):
    import torch
    import torch.distributed as dist
    from cpprcoder_b200 import api, container
    from cpprcoder_b200 import dist as rcdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("N > 1 runs under torchrun (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU path to time")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local)
    if world > 1:
        with _StdoutToStderr():
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()  # brings the communicator up (and its banner out) here

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = api.Context(local)
    env = {"world": world, "rank": rank, "local": local, "local_world": local_world, "dev": dev, "ctx": ctx,
           "barrier": barrier}
    cfg, gen, nbytes, mode, block = config_of(args, world)
    cfg["numa"] = numa
    warmup = max(args.warmup, 3)
    coder_tuple = cpu_coder()

    # ---- the headline leg: weak scaling, device resident
    leg = Leg(env, gen, nbytes * world, mode, block)
    n, n_total = leg.n, leg.n_total
    sampler = ClockSampler(local)
    for _ in range(warmup):  # warm up before the clocks are watched
        leg.step()
    sampler.start()
    main = leg.run(0, args.steps, profile=True)
    clocks = sampler.stop()
    comp_bytes = leg.used

    # ---- parity gate: payload bytes against the reference's, on every rank; N = 1 compares the
    #      whole CPU-baseline sample (its payloads are the reference's own)
    cpu_line, ref_keep = None, {}
    if world == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample, n)
        sample -= sample % block
        cpu_line = cpu_baseline(coder_tuple, leg.data[:sample], mode, block, ref_keep)
        checked = leg.parity(coder_tuple, sample // block, ref_keep)
    else:
        checked = leg.parity(coder_tuple, args.parity_blocks)
    pc = torch.tensor([checked], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(pc, op=dist.ReduceOp.SUM)
    parity_blocks = int(pc.item())

    # ---- end to end through the host-pointer C ABI, pinned host buffers
    e2e = None
    if not args.no_e2e:
        h_src = torch.from_numpy(leg.data).pin_memory()
        h_enc = torch.empty(leg.bound, dtype=torch.uint8).pin_memory()
        h_dec = torch.empty(n, dtype=torch.uint8).pin_memory()
        a_src, a_enc, a_dec = h_src.numpy(), h_enc.numpy(), h_dec.numpy()
        nb_local = leg.blk_hi - leg.blk_lo

        def step_host():
            if mode == BLKSORT:
                out = ctx.blk_encode(a_src, dst=a_enc)                 # H2D n, kernels, D2H coded
                ctx.blk_decode(out, dst=a_dec)                         # H2D coded, kernels, D2H n
                return out.size
            out = ctx.encode(mode, a_src, block, dst=a_enc)           # H2D n, kernels, D2H container
            if world > 1:  # the stitched index needs every rank's sizes: the same small collective
                offs = np.frombuffer(out[container.HEADER:container.HEADER + 8 * (nb_local + 1)], dtype=np.int64)
                rcdist.allgather_sizes(torch.from_numpy(np.diff(offs).astype(np.int32)).to(dev), n_total, block)
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
        assert bytes(a_dec[:4096]) == bytes(leg.data[:4096]) and bytes(a_dec[-4096:]) == bytes(leg.data[-4096:])
        tw = torch.tensor([wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
        ceiling_s = copy_ceiling(env, n, int(used_host), h_src, h_enc)
        e2e = {"value": n_total * args.steps / wall / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(n + used_host),
               "d2h_bytes_per_step": int(used_host + n), "ms_per_step": 1e3 * wall / args.steps,
               "api": ("b2rc_blk_encode + b2rc_blk_decode" if mode == BLKSORT else "b2rc_encode + b2rc_decode") +
                      " (host pointers, pinned)",
               "copy_ceiling_GBps": n_total / ceiling_s / 1e9, "copy_ceiling_ms_per_step": 1e3 * ceiling_s,
               "frac_of_copy_ceiling": (1e3 * ceiling_s) / (1e3 * wall / args.steps),
               "copy_ceiling_how": "the step's four copies alone (H2D input | D2H container, then H2D container | D2H "
                                   "output) with plain cudaMemcpyAsync on two streams, same pinned buffers, all ranks "
                                   "at once, max over ranks"}
        del h_src, h_enc, h_dec

    # ---- BASELINE config 3 as written: ONE stream of `nbytes` cut over the N ranks
    strong = None
    if world > 1 and mode != BLKSORT:
        del leg.enc, leg.dec
        sleg = Leg(env, gen, nbytes, mode, block)
        s = sleg.run(warmup, args.steps)
        sp = sleg.parity(coder_tuple, args.parity_blocks)
        strong = {"value": s["value"], "unit": UNIT, "global_bytes": nbytes, "ms_per_step": s["ms_per_step"],
                  "encode_GBps": s["encode_GBps"], "decode_GBps": s["decode_GBps"], "parity_blocks_checked_this_rank": sp,
                  "speedup_over_one_gpu_share": s["value"] / (main["value"] / world),
                  "efficiency": s["value"] / main["value"],
                  "compressed_ratio": s["compressed_ratio"],
                  "restart_syms": ctx.restart_for(mode, block, -(-(nbytes // block) // world)) if mode <= 3 else 0,
                  "note": "same stream as 1 GPU codes alone, blocks sharded over the ranks; efficiency = strong value / "
                          "weak value of this run (= N x the per-GPU rate on a full 1 GiB shard)"}
        del sleg

    # ---- extra legs (N = 1 default run only): the other configs' corners, device resident
    extra = None
    if world == 1 and not args.no_extra and args.workload == "zipf1g-static-64k" and not args.bytes:
        extra = {}
        x_steps = min(args.steps, 3)
        # the same containers without restart points (what a reference-shaped writer produces): one chain per block
        plain = make_ctx(api, local, B2RC_RESTART_SYMS=0)
        _, used0 = plain.encode_device(mode, leg.src, leg.enc, block)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ctx.decode_device(leg.enc, used0, leg.dec)
        ev[0].record()
        for _ in range(x_steps):
            ctx.decode_device(leg.enc, used0, leg.dec)
        ev[1].record()
        torch.cuda.synchronize()
        assert torch.equal(leg.dec[:n], leg.src)
        extra["decode_GBps_no_restart"] = n * x_steps / (ev[0].elapsed_time(ev[1]) / 1e3) / 1e9
        plain.close()
        del leg.enc, leg.dec
        for name in ("zipf1g-static-1m", "mixed-adaptive-64k"):
            g2, nb2, m2, b2 = WORKLOADS[name]
            x = Leg(env, g2, nb2, m2, b2)
            r = x.run(2, x_steps)
            pb = x.parity(coder_tuple, 64 if b2 > 65536 else 256)
            extra[name] = {"value": r["value"], "unit": UNIT, "encode_GBps": r["encode_GBps"], "decode_GBps": r["decode_GBps"],
                           "compressed_ratio": r["compressed_ratio"], "parity_blocks_checked": pb, "steps": x_steps}
            del x

    # ---- the C++ drop-in classes end to end (what a caller of RangeEncoder<> gets): tools/e2e_cpp, its own process
    e2e_cpp = None
    exe = ROOT / "tools" / "e2e_cpp"
    if world == 1 and not args.no_e2e and not args.no_extra and mode in (0, 1) and block == 65536 and gen == "zipf" and exe.exists():
        import subprocess
        try:
            r = subprocess.run([str(exe), str(n), str(min(args.steps, 3)), str(mode)], capture_output=True, text=True,
                               timeout=600, env=dict(os.environ, CPPRCODER_B200_DEVICE=str(local)))
            if r.returncode == 0:
                e2e_cpp = json.loads(r.stdout.strip().splitlines()[-1])
                assert e2e_cpp["first_mib_byte_sum"] == int(leg.data[:1 << 20].astype(np.uint64).sum()), "e2e_cpp coded another stream"
            else:
                e2e_cpp = {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as e:  # a missing compiler or binary must not cost the headline its line
            e2e_cpp = {"error": repr(e)[:300]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)"
        nb_local = leg.blk_hi - leg.blk_lo
        kavg = main["kernel_ms"]
        algo = {"histogram": n + (512 if mode == 0 else 1032) * nb_local, "ranges": n + 264 * nb_local,
                "encode": n + comp_bytes, "seams": 64 * nb_local, "scan": 12 * nb_local, "compact": 2 * comp_bytes,
                "decode": comp_bytes + n, "blk_forward": n + comp_bytes, "blk_inverse": comp_bytes + n}
        kernels = {k: {"ms": ms, "algorithmic_bytes": algo[k], "GBps": algo[k] / ms / 1e6, "hbm_frac": algo[k] / ms / 1e6 / peak}
                   for k, ms in kavg.items() if k in algo and ms > 0}
        fwd_k, inv_k = ("blk_forward", "blk_inverse") if mode == BLKSORT else ("encode", "decode")
        dom = max((k for k in kernels if k in (fwd_k, inv_k)), key=lambda k: kernels[k]["ms"]) if kernels else None
        prof = {}
        try:  # per-launch figures from the committed ncu captures: DRAM bytes, issue slots, instructions
            prof = json.loads((ROOT / "profiles" / "kernel_stats.json").read_text())
        except Exception:
            pass
        roofline = None
        if dom:
            kname = KERNEL_NAMES[mode][0 if dom == fwd_k else 1]
            st = prof.get(f"{args.workload}:{kname}", {})
            roofline = {"kernel": kname,
                        "bound": "issue" if mode != BLKSORT else "shared-memory",
                        "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                        "frac": kernels[dom]["hbm_frac"], "traffic": st.get("dram_bytes"), "peak_source": peak_src,
                        "traffic_source": st.get("source"),
                        "issue_active_pct": st.get("issue_active_pct"), "inst_per_symbol": st.get("inst_per_symbol"),
                        "note": ("shared-memory bound sort kernel (one 32 KiB block per CTA, every pass a gather and a "
                                 "scatter through shared memory); HBM fraction shown for context") if mode == BLKSORT else
                                ("coder kernels are bound by integer issue slots (serial chains of ~50-100 instructions "
                                 "per symbol), not by HBM: frac/achieved/peak are the HBM context figure the contract asks "
                                 "for, issue_active_pct and inst_per_symbol (ncu, profiles/) are the roofline that binds"),
                        "hbm_kernels": {k: {"GBps": v["GBps"], "frac": v["hbm_frac"]} for k, v in kernels.items()
                                        if k in ("histogram", "compact")}}
        line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": main["launches"], "roofline": roofline,
                "parity": "every compared payload byte-identical to the reference's", "parity_blocks_checked": parity_blocks,
                "encode_GBps": main["encode_GBps"], "decode_GBps": main["decode_GBps"],
                "compressed_ratio": comp_bytes / n, "kernels": kernels}
        if strong:
            line["strong"] = strong
            line["strong_GBps"] = strong["value"]
            line["strong_efficiency"] = strong["efficiency"]
        if extra:
            line["extra"] = extra
        if e2e_cpp:
            line["e2e_cpp"] = e2e_cpp
        if cpu_line:
            line["cpu_baseline"] = cpu_line
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
