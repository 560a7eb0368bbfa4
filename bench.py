#!/usr/bin/env python
"""bench.py — positions/s of the batched network-evaluation path (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path (GPU encoder -> 41 tcgen05 convs, the last with the head 1x1 convs fused -> head FC kernel: 43 launches) over one batch of 256 synthetic
positions on the 20-block x 256-filter net (BASELINE.json configs[1]), weights from a seeded synthetic
TF-format checkpoint (the bootstrap checkpoint is an un-fetched LFS object).  `value` = positions/s with inputs
resident in HBM; `e2e` = the same through the reference-facing Evaluate() entry (ug_eval_run_planes) with host
buffers.  `--impl reference` times the CPU restatement of the reference's TF-CPU evaluator (oracle/) instead.
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

METRIC = "nn_positions_per_s_batch256_20block"
UNIT = "positions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ugeval", choices=["ugeval", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--filters", type=int, default=256)
    ap.add_argument("--ctas", type=int, default=0, help="tower kernel cta_group (1/2); 0 = library default")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16"],
                    help="operand format of the tcgen05 tower (same MMA rate; fp16 meets the parity tolerance)")
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--sp-games", type=int, default=768, help="concurrent self-play games per GPU")
    ap.add_argument("--sp-threads", type=int, default=0, help="search threads per GPU (0 = auto)")
    ap.add_argument("--sp-playouts", type=int, default=1600, help="playouts per move (BASELINE configs[2])")
    ap.add_argument("--sp-moves", type=int, default=2, help="moves per game in the bounded self-play sample")
    ap.add_argument("--sp-preludes", default="80,1,200",
                    help="self-play samples: random legal moves played before the first search (first = headline)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the N=1 extras (small-batch latency, configs[4] sweep, encoder/head rooflines at batch 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload(batch, blocks, filters):
    return (f"batched evaluation of {batch} synthetic 19x19 positions per GPU, {blocks}-block x {filters}-filter "
            "dlcfg-19 net (BASELINE configs[1]), seeded synthetic TF-format checkpoint")


def checkpoint(blocks, filters, reference_arm=False):
    """seeded synthetic TF-format checkpoint.  reference_arm: written without the product library (pure-Python CRC over
    the index blocks only, no per-tensor CRC field) so that the reference process never maps libugeval.so"""
    from unrealgo_b200 import tfformat
    tag = "ref" if reference_arm else (os.getpid() if os.environ.get("RANK") else "x")
    d = f"/tmp/ugeval_bench_ckpt_{blocks}x{filters}_{tag}"
    meta = os.path.join(d, "bootstrap-ckpt.meta")
    if not os.path.exists(os.path.join(d, "bootstrap-ckpt.index")):
        if reference_arm:
            tfformat.NATIVE_CRC = False
        tfformat.write_synthetic_checkpoint(d, blocks, filters, seed=20260119, tensor_crc=not reference_arm)
    return meta, os.path.join(d, "bootstrap-ckpt")


class ClockSampler(threading.Thread):
    """SM clock, power and throttle reasons during the timed region (B200_PROFILING.md recipe).  NVML through
    pynvml when it is importable (a sample every ~2 ms, so even a 100 ms region gets dozens), else nvidia-smi."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []       # (sm_mhz, max_mhz, watts, set of reasons)
        self.stop_flag = False
        self.source = "nvidia-smi"

    def _nvml_loop(self):
        import pynvml as nv
        nv.nvmlInit()
        # CUDA_VISIBLE_DEVICES may renumber devices: ask torch for the UUID of the device in use
        import torch
        uuid = str(torch.cuda.get_device_properties(self.index).uuid)
        h = None
        for i in range(nv.nvmlDeviceGetCount()):
            hi = nv.nvmlDeviceGetHandleByIndex(i)
            u = nv.nvmlDeviceGetUUID(hi)
            u = u.decode() if isinstance(u, bytes) else u
            if uuid in u:
                h = hi
        if h is None:
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        self.source = "nvml"
        while not self.stop_flag:
            mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            self.rows.append((mhz, mx, watts, {n for n, bit in names if mask & bit}))
            time.sleep(0.002)

    def _smi_loop(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    reasons = {n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                                  "sw_power_cap"), parts[3:7]) if v.lower().startswith("active")}
                    self.rows.append((float(parts[0]), float(parts[1]), float(parts[2]), reasons))
            except Exception:
                pass
            time.sleep(0.1)

    def run(self):
        try:
            self._nvml_loop()
        except Exception:
            self.source = "nvidia-smi"
            self._smi_loop()

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        for r in self.rows:
            reasons |= r[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1],
                "power_w_max": max(r[2] for r in self.rows), "samples": len(self.rows),
                "reasons": sorted(reasons), "source": self.source}


def cpu_reference(args, seconds, batch):
    """times the oracle (fp32 torch-CPU interpretation of the same graph) on all host threads"""
    import numpy as np
    import torch
    from oracle.net_oracle import GraphOracle
    from unrealgo_b200 import synth
    meta, prefix = checkpoint(args.blocks, args.filters)
    torch.set_num_threads(os.cpu_count() or 1)
    orc = GraphOracle(meta, prefix)
    planes = synth.packed_to_planes(synth.synthetic_packed_positions(batch, seed=7))
    cores = torch.get_num_threads()
    orc.evaluate_reference(planes[:1])  # warm
    t0 = time.perf_counter()
    done = 0
    while True:
        orc.evaluate_reference(planes)
        done += batch
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return {"value": done / el, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{done} positions in batches of {batch} ({args.blocks}x{args.filters}), fp32 torch-CPU "
                      f"graph interpreter restating the reference's TF-CPU Session::Run; {el:.1f} s"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 16  # MAX_BATCHES: the largest batch the reference evaluator can express
    per_step = max(args.cpu_seconds / max(args.steps + args.warmup, 1), 0.5)
    import numpy as np
    import torch
    from oracle.net_oracle import GraphOracle
    from unrealgo_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)  # torchrun exports OMP_NUM_THREADS=1; this arm owns the host cores
    meta, prefix = checkpoint(args.blocks, args.filters, reference_arm=True)
    orc = GraphOracle(meta, prefix)
    planes = synth.packed_to_planes(synth.synthetic_packed_positions(batch, seed=7))
    for _ in range(args.warmup):
        orc.evaluate_reference(planes[:4])
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        ts = time.perf_counter()
        while True:
            orc.evaluate_reference(planes)
            done += batch
            if time.perf_counter() - ts >= per_step * 0.5:
                break
    el = time.perf_counter() - t0
    v = done / el
    cb = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
          "sample": f"{done} positions, batch {batch}, {args.blocks}x{args.filters}, fp32 torch-CPU restatement "
                    f"(TensorFlow and the bootstrap checkpoint are unavailable offline)"}
    emit({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * el / args.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic",
                      "config": {"workload": workload(args.batch, args.blocks, args.filters), "batch": args.batch,
                                 "blocks": args.blocks, "filters": args.filters,
                                 "sample": f"each step evaluates positions of that workload {batch} at a time (MAX_BATCHES, "
                                           "funcapproximator/DlTFNetworkEvaluator.h:15: the largest batch the reference's "
                                           "evaluator accepts) for a bounded time on all host threads"},
                      "cpu_baseline": cb,
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line of the contract, on the process's real stdout"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    global _REAL_STDOUT, METRIC
    args = parse()
    METRIC = f"nn_positions_per_s_batch{args.batch}_{args.blocks}block"  # the default is BASELINE.json's metric
    # Libraries print on stdout behind our back (NCCL writes its version line there): park the real stdout and
    # route everything else to stderr, so that stdout carries exactly the JSON line.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        return main_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    from unrealgo_b200 import PREC_BF16, PREC_FP16, DlTFNetworkEvaluator, selfplay, shard, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # one process per GPU: give every rank its own slice of the host CPUs (search threads, gather and completion
    # threads inherit it) so that ranks do not migrate over each other's caches
    affinity = None
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if world > 1 and len(cpus) >= world:
            per = len(cpus) // world
            affinity = cpus[local * per:(local + 1) * per]
            os.sched_setaffinity(0, affinity)
    except (AttributeError, OSError):
        affinity = None
    if world > 1:
        import datetime
        # a rank that dies must not leave the others in a 10-minute collective timeout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    B = args.batch
    meta, prefix = checkpoint(args.blocks, args.filters)
    # 262 positions = 5 full waves of the tower kernel (ug_eval_efficient_batch): the self-play leg batches up to that
    SP_BATCH = max(B, 262)
    ev = DlTFNetworkEvaluator(meta, device=local, max_batch=SP_BATCH,
                              precision=PREC_FP16 if args.precision == "fp16" else PREC_BF16)
    assert ev.UpdateCheckPoint(prefix), ev._err()
    if args.ctas:
        ev.set_conv_ctas(args.ctas)
    info = ev.info()

    # 8 different input batches, rotated, so no step re-reads the inputs of the previous one
    nrot = 8
    host_packed = [synth.synthetic_packed_positions(B, seed=100 + rank * 16 + i) for i in range(nrot)]
    d_pos = [torch.from_numpy(p.view(np.int32)).cuda() for p in host_packed]
    d_pol = torch.empty(B, 362, dtype=torch.float32, device="cuda")
    d_val = torch.empty(B, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    # a real (non-legacy) stream: the kernels are launched on it and the torch events below are recorded on it
    tstream = torch.cuda.Stream()
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step(i):
        ev.run_packed_device(d_pos[i % nrot].data_ptr(), B, d_pol.data_ptr(), d_val.data_ptr(), stream=stream)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush.zero_()  # evict L2 between timed iterations (outside the event bracket)
        evs[i][0].record()
        step(i)
        evs[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    assert torch.isfinite(d_pol).all() and torch.isfinite(d_val).all()

    # A.6.8 on hardware: every GPU returns the single-GPU result.  One fixed batch (the same on every rank), a digest of
    # the policy/value bytes, all ranks must agree.
    identical = None
    if world > 1:
        import hashlib
        fixed = synth.synthetic_packed_positions(64, seed=4242)
        fp, fv = ev.evaluate_packed(fixed)
        digest = int.from_bytes(hashlib.sha256(fp.tobytes() + fv.tobytes()).digest()[:7], "little")
        mine = torch.tensor([digest], dtype=torch.int64, device="cuda")
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        identical = all(int(t.item()) == digest for t in every)
        assert identical, f"rank {rank}: the GPUs disagree on a fixed 64-position batch: {[int(t.item()) for t in every]}"

    # end to end through the reference-facing entry: host int8 planes in, host double policy/value out
    planes = [synth.packed_to_planes(p) for p in host_packed[:2]]
    pol = np.empty((B, 362), np.float64)
    val = np.empty(B, np.float64)
    # the reference's EvalBuffer is one fixed allocation: page-lock the caller arrays once, as its shim does
    registered = ev.register_host(planes[0], planes[1], pol, val)
    for i in range(3):
        assert ev.Evaluate(planes[i % 2], pol, val, B)
    if world > 1:
        dist.barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        assert ev.Evaluate(planes[i % 2], pol, val, B)
    e2e_s = time.perf_counter() - t0
    # native packed host entry as a second data point
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ev.evaluate_packed(host_packed[i % nrot])
    e2e_packed_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s, e2e_packed_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s, e2e_packed_s = (float(x) for x in te.tolist())
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel (3x3 CxC tower conv), in the two regimes the part runs in.
    # burst: short event brackets on a cool GPU, the regime of the timed steps above (`value`);
    # sustained: the same brackets taken straight after seconds of back-to-back steps, at the power cap.
    conv_launches = 2 * info["blocks"]
    C = info["filters"]
    conv_flops = 2.0 * B * 361 * 9 * C * C

    def kernel_groups(iters):
        br = ev.bench(host_packed[0], iters)
        ev.set_debug_stop(0)
        stem = ev.bench(host_packed[0], iters)["ms_tower"]
        ev.set_debug_stop(-1)
        br["ms_stem"] = stem
        br["ms_conv"] = (br["ms_tower"] - stem) / conv_launches
        return br

    time.sleep(0.5)
    br_burst = kernel_groups(2)

    # sustained rate: the same step back to back for ~3 s (no L2 flush, no idle gaps) — the part then sits at its
    # power cap; this is the rate the self-play leg can at best reach
    t0 = time.perf_counter()
    sus_steps = 0
    while time.perf_counter() - t0 < 3.0:
        for i in range(32):
            step(sus_steps + i)
        torch.cuda.synchronize()
        sus_steps += 32
    sustained = world * B * sus_steps / (time.perf_counter() - t0)
    br_sus = kernel_groups(60)
    ts = torch.tensor([sustained], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MIN)  # every rank reports world * its own rate: keep the slowest
    sustained = float(ts.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_sus = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_burst = peaks.get("bf16_tflops", 1590.0)
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    traffic = None
    try:  # DRAM bytes per launch from the committed ncu capture (mean of the two tower-conv variants)
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["tower_conv_mean"]
    except Exception:
        pass

    def conv_roofline(br, regime, peak, source):
        achieved = conv_flops / (br["ms_conv"] * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "regime": regime, "peak_source": source if peaks else "fallback",
                "kernel": (f"conv3x3_v2_kernel<64> cta_group::2 (3x3 {C}->{C}, M={B * 361})" if info.get("conv_impl") == 2
                           else f"conv3x3_tc_kernel<64,{info['conv_ctas']}> (3x3 {C}->{C}, M={B * 361})"),
                "flops_per_launch": conv_flops, "ms_per_launch": br["ms_conv"], "launches_per_step": conv_launches,
                "ms_tower_conv_per_step": br["ms_conv"] * conv_launches, "traffic": traffic,
                "traffic_note": "DRAM read+write bytes per launch, mean of the two tower-conv variants, from "
                                "profiles/r02_traffic.json (ncu --set full, cold L2); tensor-bound kernel"}

    def hbm_line(kernel, batch, ms, bytes_per_pos, note):
        gbs = batch * bytes_per_pos / (ms * 1e-3) / 1e9
        return {"kernel": kernel, "bound": "hbm", "batch": batch, "us_per_launch": ms * 1e3,
                "algorithmic_bytes_per_position": bytes_per_pos, "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": gbs / hbm_peak, "note": note}

    # encoder and head kernels (HBM-bound by SURVEY 8(d)); at batch 256 they are one wave of CTAs and latency-bound
    ENC_B, HEADCONV_B = 13011, 186284                      # SURVEY 8(d) bytes per position
    HEADFC_B = 361 * (info["policy_channels"] + info["value_channels"]) * 4 + 363 * 4
    aux = [hbm_line("encode_packed_nhwc_kernel", B, br_burst["ms_encoder"], ENC_B,
                    "packed masks in + unpadded 16-bit NHWC planes out"),
           hbm_line("head_fc_cluster_kernel", B, br_burst["ms_heads"], HEADFC_B,
                    "the head stage that still exists at batch 256: fp32 head features in, policy/value out (the 1x1 head "
                    "convolutions are fused into the last tower layer, which removes SURVEY 8(d)'s 184 832-byte "
                    "activation read per position); FC weights (1.4 MB per 4-CTA cluster) not counted: L2/latency-bound")]

    # ---- N = 1 extras on rank 0: small batches, configs[4], HBM kernels at a bandwidth-saturating batch
    small_batch = configs4 = None
    if world == 1 and not args.no_extras:
        try:
            small_batch = {}
            for n in (1, 16):
                pl, po, va = planes[0][:n].copy(), np.empty((n, 362), np.float64), np.empty(n, np.float64)
                ev.register_host(pl, po, va)
                row = {}
                for name, sb in (("small_batch_kernel", 32), ("large_batch_kernel_only", 0)):
                    ev.set_small_batch(sb)
                    # these calls keep a few dozen SMs busy for a third of a millisecond each: warm up long enough for
                    # the clocks to come back from the host-bound section before, then the median of three blocks
                    for _ in range(300):
                        ev.Evaluate(pl, po, va, n)
                    reps, blocks = 200, []
                    for _ in range(3):
                        t0 = time.perf_counter()
                        for _ in range(reps):
                            ev.Evaluate(pl, po, va, n)
                        blocks.append((time.perf_counter() - t0) / reps)
                    dt = sorted(blocks)[1]
                    row[name] = {"ms_per_evaluate": dt * 1e3, "positions_per_s": n / dt,
                                 "blocks_ms": [round(b * 1e3, 4) for b in blocks]}
                ev.set_small_batch(32)
                ev.unregister_host(pl, po, va)
                small_batch[f"n{n}"] = row
            small_batch["note"] = ("DlTFNetworkEvaluator::Evaluate with host buffers at the batch sizes the reference's own "
                                   "engine issues (1 = the shipped FORCE_SINGLE_THREAD build, 16 = MAX_BATCHES): "
                                   "conv3x3_sb tower against conv3x3_v2 alone")
        except Exception as e:
            small_batch = {"error": repr(e)}
        try:
            os.environ["UGEVAL_FUSE_HEADS"] = "0"
            m1, p1 = checkpoint(1, C)
            big = DlTFNetworkEvaluator(m1, device=local, max_batch=4096,
                                       precision=PREC_FP16 if args.precision == "fp16" else PREC_BF16)
            os.environ.pop("UGEVAL_FUSE_HEADS", None)
            assert big.UpdateCheckPoint(p1), big._err()
            pk = synth.synthetic_packed_positions(512, seed=5)
            pk4096 = np.concatenate([pk] * 8)
            for nb in (B, 4096):
                bb = big.bench(pk4096[:nb], 5)
                if nb != B:
                    aux.append(hbm_line("encode_packed_nhwc_kernel", nb, bb["ms_encoder"], ENC_B,
                                        "bandwidth-saturating batch"))
                # unfused head path = head_conv_kernel + head FC kernel; the FC part alone is what the fused build times
                fc_ms = br_burst["ms_heads"] * nb / B if nb != B else br_burst["ms_heads"]
                aux.append(hbm_line("head_conv_kernel (UGEVAL_FUSE_HEADS=0) + head FC", nb, bb["ms_heads"], HEADCONV_B,
                                    "the unfused head path SURVEY 8(d) describes: reads the final activation (hi+lo 16-bit "
                                    "pair: 2x the counted bytes actually move); not on the default path at this width"
                                    + ("" if nb == B else f"; FC part alone ~{fc_ms * 1e3:.0f} us by scaling")))
            big.close()
        except Exception as e:
            os.environ.pop("UGEVAL_FUSE_HEADS", None)
            aux.append({"error": repr(e)})
        try:
            m40, p40 = checkpoint(40, 256)
            ev40 = DlTFNetworkEvaluator(m40, device=local, max_batch=1024,
                                        precision=PREC_FP16 if args.precision == "fp16" else PREC_BF16)
            assert ev40.UpdateCheckPoint(p40), ev40._err()
            i40 = ev40.info()
            pk = np.concatenate([host_packed[0], host_packed[1], host_packed[2], host_packed[3]])
            d40 = torch.from_numpy(pk.view(np.int32)).cuda()
            dp40 = torch.empty(1024, 362, dtype=torch.float32, device="cuda")
            dv40 = torch.empty(1024, dtype=torch.float32, device="cuda")
            sweep = {}
            for nb in (64, 128, 256, 512, 1024):
                for _ in range(3):
                    ev40.run_packed_device(d40.data_ptr(), nb, dp40.data_ptr(), dv40.data_ptr(), stream=stream)
                torch.cuda.synchronize()
                pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
                for a, b in pairs:
                    flush.zero_()
                    a.record()
                    ev40.run_packed_device(d40.data_ptr(), nb, dp40.data_ptr(), dv40.data_ptr(), stream=stream)
                    b.record()
                torch.cuda.synchronize()
                ms = sum(a.elapsed_time(b) for a, b in pairs) / len(pairs)
                pps = nb / (ms * 1e-3)
                sweep[str(nb)] = {"ms_per_step": ms, "positions_per_s": pps,
                                  "tensor_frac_of_burst_peak": pps * i40["flops_per_position"] / (peak_burst * 1e12)}
            configs4 = {"workload": "BASELINE configs[4]: 40-block x 256-filter tower, random-init weights, batch sweep "
                                    "(isolated steps, L2 flushed between them, CUDA events)",
                        "flops_per_position": i40["flops_per_position"], "batch512": sweep["512"], "sweep": sweep,
                        "target_60pct_positions_per_s": 0.6 * peak_burst * 1e12 / i40["flops_per_position"]}
            ev40.close()
            del d40, dp40, dv40
        except Exception as e:
            configs4 = {"error": repr(e)}

    # self-play playouts/s (BASELINE configs[2]/[3]): games sharded per GPU, no collective on the data path.
    # Samples start from positions reached by k random legal moves (k = 80: middle game, the headline; 1: opening, the
    # round-1 sample; 200: late middle game): host cost per playout and tree shape depend on the stage of the game.
    sp = None
    if not args.no_selfplay:
        ncpu = len(affinity) if affinity else (os.cpu_count() or 8) // max(world, 1)
        threads = args.sp_threads or max(2, min(12, ncpu))
        legs = {}
        for k in [int(x) for x in args.sp_preludes.split(",") if x.strip() != ""]:
            if world > 1:
                dist.barrier()
            moves = args.sp_moves if k == int(args.sp_preludes.split(",")[0]) else 1
            st, _ = selfplay(ev, games=args.sp_games, threads=threads, playouts_per_move=args.sp_playouts,
                             max_moves=moves, max_batch=SP_BATCH, timeout_us=300, reuse_tree=True,
                             prelude_moves=k, seed=1000 + rank)
            # whole-job rate = playouts of all ranks / slowest rank's time (shard.reduce_throughput: SUM, MAX)
            playouts, sp_secs = shard.reduce_throughput(st["playouts"], st["seconds"], device="cuda")
            evals, _ = shard.reduce_throughput(st["evaluations"], st["seconds"], device="cuda")
            legs[str(k)] = {"playouts_per_s": playouts / sp_secs, "evaluations_per_s": evals / sp_secs,
                            "seconds": sp_secs, "moves_per_game_sampled": moves, "mean_batch_rank0": st["mean_batch"],
                            "flush_timeouts_rank0": st["flush_timeouts"], "mean_tree_nodes_rank0": st["mean_tree_nodes"],
                            "search_busy_fraction_rank0": st["search_busy_fraction"],
                            "gpu_busy_fraction_rank0": st["gpu_busy_fraction"],
                            "host_us_per_playout_rank0": st["host_us_per_playout"],
                            "batches_rank0": st["batches"], "full_batches_rank0": st["full_batches"],
                            "wave_cut_batches_rank0": st["wave_cut_batches"],
                            "max_queue_depth_rank0": st["max_queue_depth"], "gather_sleeps_rank0": st["gather_sleeps"]}
        head = legs[str(int(args.sp_preludes.split(",")[0]))]
        sp = dict(head)
        sp.update({"unit": "playouts/s", "sample": f"headline: games started after {args.sp_preludes.split(',')[0]} random "
                                                  "legal moves (middle game); by_random_prelude_moves lists every sample",
                   "by_random_prelude_moves": legs,
                   "games_per_gpu": args.sp_games, "search_threads_per_gpu": threads, "host_cpus": os.cpu_count(),
                   "cpu_affinity_rank0": affinity,
                   "leaf_batch_cap": SP_BATCH, "playouts_per_move": args.sp_playouts, "reuse_searchtree": True,
                   "sharding": "independent games per GPU, weights replicated, no collective"})
        # one game alone (GTP-style use): one leaf in flight vs the reference's virtual-loss multi-thread search
        try:
            # one leaf in flight: nothing to batch, so no flush window (timeout 0): the rate is 1 / evaluation latency
            s1, _ = selfplay(ev, games=1, threads=1, playouts_per_move=800, max_moves=2, max_batch=SP_BATCH,
                             timeout_us=-1, reuse_tree=True, prelude_moves=80, seed=77, leaves_per_game=1)
            s64, _ = selfplay(ev, games=1, threads=1, playouts_per_move=args.sp_playouts, max_moves=3,
                              max_batch=SP_BATCH, timeout_us=300, reuse_tree=True, prelude_moves=80, seed=77,
                              leaves_per_game=64)
            sp["single_game"] = {"one_leaf_playouts_per_s": s1["playouts_per_s"],
                                 "virtual_loss_64_leaves_playouts_per_s": s64["playouts_per_s"],
                                 "virtual_loss_mean_batch": s64["mean_batch"]}
        except Exception as e:  # an extra, never part of the headline
            sp["single_game"] = {"error": repr(e)}

    if rank == 0:
        value = world * B * args.steps / (dev_ms * 1e-3)
        e2e_v = world * B * e2e_steps / e2e_s
        deviation = None
        if args.precision == "fp16":
            deviation = ("fp16 operands instead of north_star's bf16: same width, same tcgen05 kind::f16 rate; bf16's 8-bit "
                         "mantissa misses the stated 1e-3 / 5e-3 tolerance on the 20-block net (DESIGN.md section 4)")
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "bf16", "data": "synthetic",
            "precision_deviation": deviation,
            "config": {"workload": workload(B, info["blocks"], C),
                       "batch": B, "blocks": info["blocks"], "filters": C, "conv_cta_group": info["conv_ctas"],
                       "conv_impl": info.get("conv_impl"),
                       "arithmetic": f"{args.precision} operands (tcgen05 kind::f16), fp32 accumulation in TMEM, "
                                     "fp32 bias/residual epilogue, residual stream carried as hi+lo 16-bit pair",
                       "l2": "flushed between timed steps (256 MiB write outside the event bracket); 8 input batches rotated",
                       "flops_per_position": info["flops_per_position"],
                       "regime_of_value": "burst: isolated steps on a cool GPU; sustained_3s_positions_per_s is the "
                                          "steady-state (power-capped) rate of the same step",
                       "tensor_frac_of_burst_peak_whole_step": value / world * info["flops_per_position"] / (peak_burst * 1e12),
                       "tensor_frac_of_sustained_peak_whole_step_sustained":
                           sustained / world * info["flops_per_position"] / (peak_sus * 1e12),
                       "e2e_packed_positions_per_s": world * B * e2e_steps / e2e_packed_s,
                       "sustained_3s_positions_per_s": sustained,
                       "kernel_group_ms": {"burst": {k: br_burst[k] for k in ("ms_encoder", "ms_tower", "ms_heads", "ms_stem")},
                                           "sustained": {k: br_sus[k] for k in ("ms_encoder", "ms_tower", "ms_heads", "ms_stem")}}},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": B * 17 * 361, "d2h_bytes_per_step": B * 363 * (8 if registered == 4 else 4),
                    "api": "DlTFNetworkEvaluator::Evaluate (ug_eval_run_planes): host int8 planes -> host double policy/value",
                    "host_buffers": f"{registered}/4 caller arrays page-locked once with ug_eval_register_host"},
            "gpu_launches": br_burst["launches_per_pass"] * args.steps,
            "roofline": conv_roofline(br_burst, "burst", peak_burst,
                                      "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed in short brackets, as `value`)"),
            "roofline_sustained": conv_roofline(br_sus, "sustained", peak_sus,
                                                "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed straight after "
                                                "3 s of back-to-back steps, at the power cap)"),
            "roofline_aux": aux,
        }
        if identical is not None:
            out["multi_gpu_identical"] = identical
        if small_batch is not None:
            out["config"]["small_batch"] = small_batch
        if configs4 is not None:
            out["configs4"] = configs4
        if sp is not None:
            out["selfplay"] = sp
        if not args.no_cpu_baseline and world == 1:
            try:
                out["cpu_baseline"] = cpu_reference(args, args.cpu_seconds, 16)
            except Exception as e:  # the baseline is reported beside the result, never part of it
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"failed: {e!r}"}
        emit(out)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
