#!/usr/bin/env python3
"""bench.py -- bootstrapped HomNAND gates/s on N B200s (BASELINE.json metric), one process per GPU.

A "step" = one pass of the hot path (gate pre-combination -> 635 x CMUX blind rotation -> sample extract -> LWE key
switch) over one batch of 1024 independent NAND gates per GPU (BASELINE.json configs[1]); weak scaling: every rank
owns its own 1024-gate shard, keys are replicated once by an NCCL broadcast, no collective per gate.

  python bench.py [--gpus N] [--steps K] [--warmup W]          GPU arm (this repo's CUDA path through the C ABI)
  python bench.py --impl reference ...                         reference arm: the reference's CPU gate path (its own FFT
                                                               library oracle/_ref + restated glue) on the host cores

One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED0001
BATCH = 1024                     # gates per GPU per step (configs[1])
N_LWE, N_POLY = 635, 1024
CT_WORDS = N_LWE + 1
NROT = 32                        # distinct input batches rotated through: 32 x 5.2 MB = 167 MB > 126 MB L2
# algorithmic figures (SURVEY.md section 8d / DESIGN.md section "Roofline")
BK_BYTES_ALGO = 62_423_040       # 635 x 12 x 1024 coefficients x 8 B (SURVEY's single-modulus basis)
KSK_BYTES = 1024 * 8 * 3 * 636 * 4
MODMUL_PER_GATE = 33.81e6        # 635 x (8 x 5120 butterflies + 12288 MACs), single-modulus basis
WORKLOAD = ("batched 1024 independent HomNAND gates per GPU (BASELINE configs[1]), default TFHE parameters "
            "n=635 N=1024 l=3 Bg=64 t=8 basebit=2, decomposition mask 0x02084000 (reference-faithful)")


def design_figures(key_slices):
    """Per-gate work of the shipped arithmetic (DESIGN.md sections 2 and 5): S key slices -> 6 forward + 2S inverse transforms of
    5120 Shoup butterflies, 12 S x 1024 wide multiply-accumulates and 2 S x 1024 Montgomery reductions per CMUX, 635 CMUX per gate.
    FMA-heavy issue slots: butterfly 4 (IMAD.HI 2 + 2 IMAD), IMAD.WIDE 2.5, reduction 3 (measured weights, profiles/intpipe_r01.json)."""
    S = key_slices
    transforms = 6 + 2 * S
    return {"transforms_per_cmux": transforms, "butterflies_per_gate": 635 * transforms * 5120,
            "fma_slots_per_gate": 635 * (transforms * 5120 * 4 + 12 * S * 1024 * 2.5 + 2 * S * 1024 * 3),
            "bk_bytes_device": 635 * 12 * S * 1024 * 4}


def fft64_figures():
    """Per-gate work of the FFT64 mode (fft64.cuh; DESIGN.md sections 2 and 5), counted per lane from the code and confirmed by ncu
    (4429 FP64 warp instructions per gate and CMUX): a forward transform of 512 complex points = 9 stages x 8 butterflies x 6
    DFMA-class operations = 432 per lane + 16 (stage 4 is computed output by output on the load side of the transpose; the digits
    enter through I2F.F64.S8, not on the FP64 pipe's DFMA count) = 448, an inverse =
    482 per lane (trivial twiddles in its first stages, untwist and rounding), one spectrum x key multiply-accumulate = 64 per
    lane; per CMUX 6 forward + 2 inverse + 12 multiply-accumulates.
    Shared-memory bytes per gate and CMUX: 8 transposes x 16 KB, 96 KB of key read from the ring, 60 KB of per-lane twiddle rows,
    22 KB for the source words (rotated reads, digit byte planes), 16 KB accumulator update."""
    fwd, inv, mac = 448 * 32, 482 * 32, 64 * 32
    per_cmux = 6 * fwd + 2 * inv + 12 * mac
    smem = (8 * 16 + 96 + 60 + 22 + 16) * 1024
    # issue slots: a DFMA-class instruction keeps the issue port for two cycles (profiles/r02_dfma_mix.json); the other 3113 warp
    # instructions of a gate and CMUX (ncu, profiles/r02_ncu_blind_rotate_f64_latest.txt) take one each
    issue = 2 * per_cmux / 32 + 3113
    return {"transforms_per_cmux": 8, "fp64_ops_per_gate": 635 * per_cmux, "smem_bytes_per_gate": 635 * smem,
            "issue_cycles_per_gate": 635 * issue, "bk_bytes_device": 635 * 12 * 512 * 16}


def workload_config(world):
    """the keys both arms share (the driver compares them)"""
    return {"workload": WORKLOAD, "batch_per_gpu": BATCH, "global_batch": BATCH * world}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one blind_rotate_t2_kernel launch from the committed `ncu --set full`
    summary (profiles/r02_ncu_blind_rotate_t2_latest.txt), scaled from the profiled batch to 1024 gates (the key is read once
    per wave, the ciphertext traffic is per gate); None when the summary is missing."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_blind_rotate_f64_latest.txt")
    try:
        tot = 0.0
        for line in open(p):
            for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
                if line.startswith(key):
                    unit = line[line.index("[") + 1:line.index("]")].lower()
                    val = float(line.split("=")[1])
                    tot += val * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        return tot or None
    except Exception:
        return None


def fp64_peak():
    """DFMA issue rate measured on this pool's B200 by tools/microbench/intpipe.cu (profiles/intpipe_r01b.json: same rate as IMAD)."""
    p = os.path.join(ROOT, "profiles", "intpipe_r01b.json")
    try:
        return json.load(open(p))["dfma"]["Gops_per_s"] * 1e9, "measured (profiles/intpipe_r01b.json, dfma)"
    except Exception:
        return 148 * 64 * 1.965e9, "nominal 148 SM x 64 lanes x 1.965 GHz"


def int_peak():
    """IMAD issue-slot peak measured on this pool's B200 by tools/microbench/intpipe.cu (profiles/intpipe_r01.json)."""
    p = os.path.join(ROOT, "profiles", "intpipe_r01.json")
    try:
        return json.load(open(p))["imad_lo"]["Gops_per_s"] * 1e9, "measured (profiles/intpipe_r01.json)"
    except Exception:
        return 148 * 64 * 1.965e9, "nominal 148 SM x 64 lanes x 1.965 GHz"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def cpu_reference_run(nthreads, gates_per_thread, keys=None):
    """Time the reference-equivalent CPU gate path (reference FFT library + restated glue) on `nthreads` host threads."""
    from oracle import oracle as O
    O.lib()
    if not O.ref_init():
        raise RuntimeError("oracle/_ref/libspqlios_ref.so missing (build it where /root/reference exists: make -C oracle ref)")
    K = keys or O.Keys(SEED)
    B = nthreads * gates_per_thread
    rng = np.random.default_rng(SEED + 1)
    x = rng.integers(0, 2, B).astype(np.uint8)
    y = rng.integers(0, 2, B).astype(np.uint8)
    c0, c1 = K.encrypt(x, 0), K.encrypt(y, 10_000_000)
    K.fourier_handle()
    O.bench_ref_gates(K, O.NAND, c0[:nthreads], c1[:nthreads], nthreads)  # warm-up: one gate per thread
    secs, out = O.bench_ref_gates(K, O.NAND, c0, c1, nthreads)
    ok = bool(np.array_equal(K.decrypt(out), 1 - (x & y)))
    return B / secs, secs, B, ok, K


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nthreads = os.cpu_count() or 1
    gpt = 2  # gates per thread per step: bounded sample of the 1024-gate batch
    from oracle import oracle as O
    K = None
    vals = []
    for it in range(args.warmup + args.steps):
        gps, secs, B, ok, K = cpu_reference_run(nthreads, gpt, K)
        if it >= args.warmup:
            vals.append((gps, secs))
        if not ok:
            print(json.dumps({"impl": "reference", "error": "reference CPU path decrypted a wrong bit"}))
            return 1
    total_gates = nthreads * gpt * args.steps
    total_secs = sum(s for _, s in vals)
    v = total_gates / total_secs
    line = {
        "impl": "reference", "metric": "bootstrapped HomNAND gates/sec", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(max(1, args.gpus)), sample_gates_per_step=nthreads * gpt,
                       note="each step is a bounded sample of the 1024-gate batch (CPU time per gate does not depend on the batch)"),
        "cpu_baseline": {"value": v, "unit": "gates/s", "cores": nthreads, "kind": "port",
                         "sample": f"{nthreads * gpt} NAND gates per step on {nthreads} threads; reference's own spqlios FFT "
                                   f"(oracle/_ref, compiled from /root/reference) + C restatement of the Rust glue (Rust toolchain absent)"},
        "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_extras(eng, R, K, torch, dev, stream, dx, dy, s0, rank, world, dist):
    """BASELINE.json configs 3, 4 and 5 as extra keys of the line (the headline stays configs[1]): device-resident, CUDA events,
    best of 3 after a warm-up.  Every rank calls this (config 5 is a collective measurement); rank 0 keeps the result."""
    import rustfhe_b200.circuit as Cq
    out = {}

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best

    S = max(2, eng.stats()["key_slices"])    # the step-level entry points run the NTT (two slices in FFT64 mode)
    fig = design_figures(S)
    p_int, _ = int_peak()
    slots_xp = fig["fma_slots_per_gate"] / 635.0                            # one external product = one CMUX step
    slots_pm = (1 + 2 * S) * 5120 * 4 + S * 1024 * (2.5 + 3)                # 1 + S forward, S inverse transforms, S x 1024 MAC + REDC
    if rank == 0:
        # ---- config 3: negacyclic poly-mul / external product at N = 1024, batch 65536 (the top of the sweep; the whole
        #      sweep is tools/sweep_config3.py) ----
        g = torch.Generator(device=dev).manual_seed(7)
        B3 = 65536
        u32 = lambda *shape: torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=dev, generator=g).to(torch.int32)
        a = u32(B3, 1024)
        d = torch.randint(-32, 32, (B3, 1024), dtype=torch.int32, device=dev, generator=g)
        o = torch.empty_like(a)
        t = timed(lambda: eng.negacyclic_mul_batch_device(a.data_ptr(), d.data_ptr(), o.data_ptr(), B3, stream.cuda_stream))
        if eng.stats()["key_slices"] == 1:   # FFT64: polymul_f64_kernel, two forward + one inverse transform per product, I/O straight from HBM
            issue_pm = 2 * 1468 + 660   # ncu (profiles/r02_ncu_polymul_f64.txt): 1468 FP64 (2 x 448 + 482 + 96 per lane) + 660 other warp instructions per product
            out["config3_negacyclic_mul"] = {"batch": B3, "products_per_s": B3 / t, "issue_roofline_frac": B3 / t * issue_pm / (148 * 4 * 1.965e9),
                                             "hbm_gbs": B3 / t * 3 * 4096 / 1e9, "note": "FFT64 arithmetic, one product per warp"}
        else:
            out["config3_negacyclic_mul"] = {"batch": B3, "products_per_s": B3 / t, "int_roofline_frac": B3 / t * slots_pm / p_int}
        del a, d, o
        trl = u32(B3, 2, 1024)
        res = torch.empty_like(trl)
        trg1 = u32(1, 6, 2, 1024)
        t = timed(lambda: eng.external_product_batch_device(trg1.data_ptr(), 1, trl.data_ptr(), res.data_ptr(), B3, stream.cuda_stream))
        if eng.stats()["key_slices"] == 1:   # FFT64: the persistent external_product_f64_kernel, bound by issue slots like the gate kernel
            issue_xp = 2 * 4415 + 2620   # ncu (profiles/r02_ncu_external_product_f64.txt): FP64 + other warp instructions per product
            out["config3_external_product_shared_trgsw"] = {"batch": B3, "products_per_s": B3 / t, "issue_roofline_frac": B3 / t * issue_xp / (148 * 4 * 1.965e9),
                                                            "hbm_gbs": B3 / t * 2 * 8192 / 1e9,
                                                            "note": "FFT64 arithmetic, one product per warp, the TRGSW transformed inside the call and kept on the key ring"}
        else:
            out["config3_external_product_shared_trgsw"] = {"batch": B3, "products_per_s": B3 / t, "int_roofline_frac": B3 / t * slots_xp / p_int,
                                                            "note": "the one TRGSW is transformed inside the call (36 of 786 k transforms)"}
        Bp = 4096
        trgB = u32(Bp, 6, 2, 1024)
        for _ in range(4):   # every work slot of the ring grows its scratch (604 MB of transformed TRGSWs) on its first use
            eng.external_product_batch_device(trgB.data_ptr(), Bp, trl.data_ptr(), res.data_ptr(), Bp, stream.cuda_stream)
        t = timed(lambda: eng.external_product_batch_device(trgB.data_ptr(), Bp, trl.data_ptr(), res.data_ptr(), Bp, stream.cuda_stream))
        out["config3_external_product_per_item_trgsw"] = {"batch": Bp, "products_per_s": Bp / t,
                                                          "note": "includes the key transforms of every item's TRGSW"}
        del trl, res, trg1, trgB
        # ---- config 4: 32-bit adders on encrypted operands, level-synchronous, one launch pair per level ----
        r = np.random.default_rng(SEED + 2)
        x, y = int(r.integers(0, 2 ** 32)), int(r.integers(0, 2 ** 32))
        bits = np.array([(x >> i) & 1 for i in range(32)] + [(y >> i) & 1 for i in range(32)], np.uint8)
        cts = R.Cryptor.encrypto(R.TLWE, s0, bits, seed=SEED + 300, ct_index0=0)
        for name, nl in (("ripple_carry_nand", Cq.ripple_carry_adder(32)), ("kogge_stone_native_gates", Cq.prefix_adder(32))):
            dc = Cq.DeviceCircuit(eng, nl)
            dc.run(cts)
            t0 = time.perf_counter()
            res = dc.run(cts)
            secs = time.perf_counter() - t0
            got = R.Cryptor.decrypto(R.TLWE, s0, res)
            out[f"config4_adder32_{name}"] = {"seconds": secs, "gates": dc.gates, "levels": dc.levels,
                                              "correct": bool(sum(int(b) << i for i, b in enumerate(got)) == x + y)}
            dc.close()
    # ---- config 5: 2^16 random-bit NAND gates, STRONG scaling: the batch is split over the ranks ----
    total = 1 << 16
    mine = total // world
    reps = -(-mine // dx.shape[0])
    bx = dx.repeat(reps, 1)[:mine].contiguous()
    by = dy.repeat(reps, 1)[:mine].contiguous()
    bo = torch.empty_like(bx)
    eng.reserve(mine)
    eng.set_batch_overlap(0)

    def sweep():
        eng.gate_batch_device(K.NAND, bx.data_ptr(), by.data_ptr(), bo.data_ptr(), mine, stream.cuda_stream)

    sweep(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); sweep(); e1.record(stream)
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    eng.set_batch_overlap(-1)
    out["config5_2pow16_gates_strong"] = {"gates": total, "n_gpus": world, "gates_per_s": total / (float(tt[0]) * 1e-3), "ms": float(tt[0])}
    # ---- config 4 on all GPUs of the run: ONE process (rank 0) drives them as a device group through the C ABI; wide levels are
    # sharded and their outputs exchanged over NCCL, levels narrower than one wave of the latency kernel are replicated ----
    cpu_wait = dist.new_group(backend="gloo") if world > 1 else None   # the other ranks wait on the CPU: an NCCL barrier would spin on their GPUs
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=cpu_wait)
    if rank == 0:
        try:
            g = R.DeviceGroup(list(range(world)))
            try:
                sk = R.SecretKeys.generate(SEED)
                g.keygen(SEED, sk.s_key_tlwelv0, sk.s_key_tlwelv1)
                r = np.random.default_rng(SEED + 4)
                for name, nl, k in (("ripple_carry_nand", Cq.ripple_carry_adder(32), 1), ("kogge_stone_x64", Cq.side_by_side(Cq.prefix_adder(32), 64), 64)):
                    xs = [int(v) for v in r.integers(0, 2 ** 32, k, dtype=np.uint64)]
                    ys = [int(v) for v in r.integers(0, 2 ** 32, k, dtype=np.uint64)]
                    bits = np.array([(v >> i) & 1 for x, y in zip(xs, ys) for v in (x, y) for i in range(32)], np.uint8)
                    cts = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, bits, seed=SEED + 400, ct_index0=0)
                    gc = Cq.GroupCircuit(g, nl)
                    gc.run(cts)
                    t0 = time.perf_counter()
                    res = gc.run(cts)
                    secs = time.perf_counter() - t0
                    got = R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, res).reshape(k, 33)
                    ok = all(sum(int(b) << i for i, b in enumerate(row)) == x + y for row, x, y in zip(got, xs, ys))
                    out[f"config4_group_{name}"] = {"n_gpus": world, "seconds": secs, "gates": gc.gates, "levels": gc.levels,
                                                    "gates_per_s": gc.gates / secs, "correct": bool(ok), **gc.last}
                    gc.close()
            finally:
                g.close()
        except Exception as ex:
            out["config4_group_error"] = str(ex)
    if world > 1:
        dist.barrier(group=cpu_wait)
    return out


def run_gpu(args):
    # keep stdout for the ONE JSON line: libraries that print to fd 1 (NCCL's "NCCL version ..." banner) go to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import rustfhe_b200 as R
    from rustfhe_b200 import _capi as K

    # ---- keys: generated once on rank 0 ON THE DEVICE (tfhe_b200_keygen_device), replicated with ONE NCCL broadcast each;
    #      the other ranks transform the received torus-domain bootstrapping key locally ----
    t_key0 = time.time()
    eng = R.DeviceEngine(local)
    eng.reserve(BATCH)
    stream = torch.cuda.current_stream()
    sk = R.SecretKeys.generate(SEED)            # seeded: every rank derives the same secret keys (needed for the decrypt checks)
    s0 = sk.s_key_tlwelv0
    if world == 1:
        eng.keygen_device(SEED, sk.s_key_tlwelv0, sk.s_key_tlwelv1)
    else:
        bk_d = torch.empty(K.BK_WORDS, dtype=torch.int32, device=dev)
        ksk_d = torch.empty(K.KSK_WORDS, dtype=torch.int32, device=dev)
        if rank == 0:
            eng.keygen_device(SEED, sk.s_key_tlwelv0, sk.s_key_tlwelv1)
            eng.export_bk_device(bk_d.data_ptr(), stream.cuda_stream)
            eng.export_ksk_device(ksk_d.data_ptr(), stream.cuda_stream)
        dist.broadcast(bk_d, 0)
        dist.broadcast(ksk_d, 0)
        if rank != 0:
            eng.load_ksk_device(ksk_d.data_ptr(), stream.cuda_stream)
            eng.load_bk_device(bk_d.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        del bk_d, ksk_d
    torch.cuda.synchronize()
    key_s = time.time() - t_key0

    # ---- synthetic inputs: NROT distinct batches of random-bit encryptions per rank, encrypted on the device, resident in
    #      HBM; a pinned host copy feeds the end-to-end leg ----
    rng = np.random.default_rng(SEED + 17 * rank)
    nct = NROT * BATCH
    bx = rng.integers(0, 2, nct).astype(np.uint8)
    by = rng.integers(0, 2, nct).astype(np.uint8)
    dx = torch.empty((nct, CT_WORDS), dtype=torch.int32, device=dev)
    dy = torch.empty((nct, CT_WORDS), dtype=torch.int32, device=dev)
    dbx, dby = torch.from_numpy(bx).to(dev), torch.from_numpy(by).to(dev)
    eng.encrypt_bits_device(SEED + 100 + rank, 0, s0, dbx.data_ptr(), nct, dx.data_ptr(), stream.cuda_stream)
    eng.encrypt_bits_device(SEED + 200 + rank, 0, s0, dby.data_ptr(), nct, dy.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    hx = torch.empty((nct, CT_WORDS), dtype=torch.int32).pin_memory()
    hy = torch.empty((nct, CT_WORDS), dtype=torch.int32).pin_memory()
    hx.copy_(dx); hy.copy_(dy)
    dout = torch.empty((BATCH, CT_WORDS), dtype=torch.int32, device=dev)

    NSTREAMS = max(1, int(os.environ.get("BENCH_STREAMS", "2")))
    streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(NSTREAMS - 1)]   # consecutive steps rotate over the streams:
    douts = [dout] + [torch.empty_like(dout) for _ in range(NSTREAMS - 1)]               # independent batches overlap on the device

    def step_device(it, st=None, o_buf=None):
        o = (it % NROT) * BATCH
        st = st or stream
        eng.gate_batch_device(K.NAND, dx[o:o + BATCH].data_ptr(), dy[o:o + BATCH].data_ptr(), (o_buf if o_buf is not None else dout).data_ptr(),
                              BATCH, st.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()

    for it in range(args.warmup):
        step_device(it)
    torch.cuda.synchronize()
    # correctness of the last warm-up batch (decrypt all 1024 outputs)
    o = ((args.warmup - 1) % NROT) * BATCH if args.warmup > 0 else 0
    if args.warmup == 0:
        step_device(0); torch.cuda.synchronize()
    got = R.Cryptor.decrypto(R.TLWE, s0, dout.cpu().numpy().view(np.uint32))
    wrong = int((got != (1 - (bx[o:o + BATCH] & by[o:o + BATCH]))).sum())

    # ---- serial region (one stream, K steps): isolated per-launch kernel durations for the roofline ----
    eng.reset_stats()
    s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    s0e.record(stream)
    for it in range(args.steps):
        step_device(args.warmup + it)
    s1e.record(stream)
    torch.cuda.synchronize(); barrier()
    serial_ms = s0e.elapsed_time(s1e)
    st = eng.stats()

    # ---- timed region: exactly K steps, device-resident inputs, steps alternate between two streams so the tail of one
    #      batch's blind rotation and its key switch run under the head of the next batch (independent batches) ----
    launches0 = eng.stats()["kernel_launches"]
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join = torch.cuda.Event()
    barrier(); torch.cuda.synchronize()
    if NSTREAMS > 1:
        eng.set_batch_overlap(1)   # batches are kept in flight on several streams: cut every one of them into full CTAs
    e0.record(stream)
    for s_ in streams[1:]:
        s_.wait_event(e0)
    for it in range(args.steps):
        step_device(args.warmup + it, streams[it % NSTREAMS], douts[it % NSTREAMS])
    for s_ in streams[1:]:
        join = torch.cuda.Event()
        join.record(s_)
        stream.wait_event(join)
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.stats()["kernel_launches"] - launches0
    eng.set_batch_overlap(-1)
    o = ((args.warmup + args.steps - 1) % NROT) * BATCH
    got = R.Cryptor.decrypto(R.TLWE, s0, douts[(args.steps - 1) % NSTREAMS].cpu().numpy().view(np.uint32))
    wrong += int((got != (1 - (bx[o:o + BATCH] & by[o:o + BATCH]))).sum())

    # ---- B = 1 latency (SURVEY 8d "latency metric"): one gate per call, median of 30 after 3 warm-ups, CUDA events ----
    lat = []
    for rep in range(33):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        eng.gate_batch_device(K.NAND, dx[rep:rep + 1].data_ptr(), dy[rep:rep + 1].data_ptr(), dout.data_ptr(), 1, stream.cuda_stream)
        a1.record(stream)
        a1.synchronize()
        if rep >= 3:
            lat.append(a0.elapsed_time(a1) * 1e3)
    latency_us = float(np.median(lat))

    # ---- e2e: the same K steps through the host-buffer C ABI (tfhe_b200_gate_batch_async + tfhe_b200_sync), pinned host
    #      memory, H2D of both input batches and D2H of the output batch inside the timed region, every step ----
    hx_np, hy_np = hx.numpy().view(np.uint32), hy.numpy().view(np.uint32)
    NOUT = 4
    houts = [torch.empty((BATCH, CT_WORDS), dtype=torch.int32).pin_memory() for _ in range(NOUT)]
    houts_np = [h.numpy().view(np.uint32) for h in houts]
    lib = K.lib()

    def step_host(it):
        o = (it % NROT) * BATCH
        rc = lib.tfhe_b200_gate_batch_async(eng._ctx, K.NAND, K.ptr(hx_np[o:o + BATCH]), K.ptr(hy_np[o:o + BATCH]),
                                            K.ptr(houts_np[it % NOUT]), BATCH)
        if rc:
            raise RuntimeError(lib.tfhe_b200_last_error(eng._ctx))

    step_host(0); eng.sync()
    eng.set_batch_overlap(1)       # the asynchronous calls below keep up to four batches in flight
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(args.steps):
        step_host(args.warmup + it)
    eng.sync()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if sampler else None
    last = args.warmup + args.steps - 1
    got = R.Cryptor.decrypto(R.TLWE, s0, houts_np[last % NOUT])
    o = (last % NROT) * BATCH
    wrong += int((got != (1 - (bx[o:o + BATCH] & by[o:o + BATCH]))).sum())

    extras = None
    if not args.no_extras:
        try:
            extras = run_extras(eng, R, K, torch, dev, stream, dx, dy, s0, rank, world, dist)
        except Exception as ex:   # the headline numbers stand on their own
            extras = {"error": str(ex)}

    # max over ranks
    t = torch.tensor([ms, e2e_s * 1e3, float(wrong), serial_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, wrong, serial_ms = float(t[0]), float(t[1]), int(t[2]), float(t[3])

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        p_int, p_int_src = int_peak()
        fft64 = st["key_slices"] == 1
        fig = fft64_figures() if fft64 else design_figures(st["key_slices"])
        BK_BYTES_DEVICE = fig["bk_bytes_device"]
        kname = "blind_rotate_f64_kernel" if fft64 else "blind_rotate_t2_kernel<6,1>" if st["key_slices"] == 2 else "blind_rotate_kernel<4,false,1,3>"
        gates = BATCH * world * args.steps
        value = gates / (ms * 1e-3)
        br_ms, ks_ms = st["avg_blind_rotate_ms"], st["avg_keyswitch_ms"]
        # dominant kernel = blind_rotate_kernel: algorithmic bytes per launch = BK once + ciphertext I/O of the batch
        algo_bytes = BK_BYTES_ALGO + BATCH * (2 * CT_WORDS * 4 + CT_WORDS * 4 + N_POLY * 2)
        achieved = algo_bytes / (br_ms * 1e-3) / 1e9 if br_ms > 0 else 0.0
        per_gpu_gps_kernel = BATCH / (br_ms * 1e-3) if br_ms > 0 else 0.0
        line = {
            "metric": "bootstrapped HomNAND gates/sec", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if fft64 else "u32", "data": "synthetic",
            "config": dict(workload_config(world),
                       arithmetic="torus words mod 2^32; " + (
                           "negacyclic f64 complex transform of the folded polynomial (512 points), products rounded to the exact integer "
                           "(6 + 2 transforms per CMUX); results bit-identical to the exact NTT modes" if fft64 else
                           "exact negacyclic NTT over the 29-bit prime 536856577, " +
                           ("two 16-bit key slices (6 + 4 transforms per CMUX)" if st["key_slices"] == 2 else "three 11-bit key slices (6 + 6 transforms per CMUX)")),
                       parallelism=f"dp{world} (independent gate shards, keys replicated)",
                       l2=f"inputs rotate over {NROT} batches ({NROT * BATCH * 2 * CT_WORDS * 4 / 1e6:.0f} MB) > L2; keys "
                          f"{(BK_BYTES_DEVICE + KSK_BYTES) / 1e6:.0f} MB ~ L2; no explicit flush",
                       gates_per_cta=st["gates_per_cta"], streams=NSTREAMS, key_slices=st["key_slices"]),
            "value_serial": gates / (serial_ms * 1e-3),
            "latency_us_per_gate_amortised": 1e3 * ms / args.steps / BATCH,
            "latency_us_single_gate": latency_us,
            "wrong_bits": wrong,
            "e2e": {"value": gates / (e2e_ms * 1e-3), "unit": "gates/s", "h2d_bytes_per_step": 2 * BATCH * CT_WORDS * 4,
                    "d2h_bytes_per_step": BATCH * CT_WORDS * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "tfhe_b200_gate_batch_async x K + tfhe_b200_sync (host pointers, pinned)"},
            "gpu_launches": int(launches),
            "kernels": {"blind_rotate_ms": br_ms, "keyswitch_ms": ks_ms, "timed_launches": st["timed_launches"],
                        "note": "isolated per-launch durations (CUDA events on the launching stream) from a serial K-step region of "
                                "this same run; in the timed region consecutive batches overlap on two streams"},
            "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                         "frac": achieved / peaks["hbm_gbs"], "peak_kind": f"{peak_kind} (MEASURED_PEAKS.json hbm_gbs)",
                         "algorithmic_bytes_per_launch": algo_bytes, "traffic": ncu_traffic(),
                         "note": "HBM is NOT the binding roof of this kernel (key bytes are read once per 1024-gate launch); "
                                 "the binding roofs are on-chip, see " + ("issue_roofline, smem_roofline and fp64_roofline" if fft64 else "int_roofline")},
            "clocks": clocks,
            "key_setup_s": key_s,
            "extras": extras,
        }
        if fft64:
            p64, p64_src = fp64_peak()
            smem_peak = 148 * 128 * (peaks.get("sm_max_mhz", 1965.0) * 1e6)    # 128 B per clock and SM (B300_MICROARCH.md, LDS/STS)
            line["fp64_roofline"] = {"bound": "FP64 pipe (DFMA issue slots)", "kernel": kname, "achieved": per_gpu_gps_kernel * fig["fp64_ops_per_gate"] / 1e12,
                                     "peak": p64 / 1e12, "unit": "T DFMA-class ops/s", "frac": per_gpu_gps_kernel * fig["fp64_ops_per_gate"] / p64,
                                     "peak_kind": p64_src, "ops_per_gate": fig["fp64_ops_per_gate"], "roof_gates_per_s": p64 / fig["fp64_ops_per_gate"],
                                     "note": "DFMA / DADD / DMUL, one issue slot each; ncu: sm__pipe_fp64_cycles_active"}
            line["smem_roofline"] = {"bound": "shared-memory bandwidth", "kernel": kname,
                                     "achieved": per_gpu_gps_kernel * fig["smem_bytes_per_gate"] / 1e9, "peak": smem_peak / 1e9, "unit": "GB/s",
                                     "frac": per_gpu_gps_kernel * fig["smem_bytes_per_gate"] / smem_peak, "peak_kind": "148 SM x 128 B/clk x SM clock",
                                     "bytes_per_gate": fig["smem_bytes_per_gate"], "roof_gates_per_s": smem_peak / fig["smem_bytes_per_gate"],
                                     "note": "algorithmic shared-memory bytes of the design (transposes, key ring reads, per-lane twiddles, source words, "
                                             "accumulator); ncu: l1tex__data_pipe_lsu_wavefronts_mem_shared"}
            issue_peak = 148 * 4 * (peaks.get("sm_max_mhz", 1965.0) * 1e6)      # one warp instruction per scheduler and clock
            line["issue_roofline"] = {"bound": "instruction issue slots (the binding roof: a DFMA holds the issue port for two cycles)", "kernel": kname,
                                      "achieved": per_gpu_gps_kernel * fig["issue_cycles_per_gate"] / 1e9, "peak": issue_peak / 1e9, "unit": "G issue cycles/s",
                                      "frac": per_gpu_gps_kernel * fig["issue_cycles_per_gate"] / issue_peak, "peak_kind": "148 SM x 4 schedulers x SM clock",
                                      "issue_cycles_per_gate": fig["issue_cycles_per_gate"], "roof_gates_per_s": issue_peak / fig["issue_cycles_per_gate"],
                                      "note": "2 x 4420 FP64 + 3113 other warp instructions per gate and CMUX; measured DFMA + FFMA mixes add up instead of "
                                              "overlapping (profiles/r02_dfma_mix.json), and a third warp per scheduler does not raise the issue rate"}
            line["int_roofline"] = dict(line["fp64_roofline"], note="FFT64 mode: the arithmetic runs on the FP64 pipe, whose issue rate equals the IMAD rate "
                                        "(profiles/intpipe_r01b.json); see fp64_roofline / smem_roofline")
        else:
            FMA_SLOTS_PER_GATE, BUTTERFLIES_PER_GATE = fig["fma_slots_per_gate"], fig["butterflies_per_gate"]
            line["int_roofline"] = {"bound": "integer FMA pipe (IMAD issue slots)", "kernel": kname,
                                    "achieved": per_gpu_gps_kernel * FMA_SLOTS_PER_GATE / 1e12, "peak": p_int / 1e12, "unit": "T IMAD-slots/s",
                                    "frac": per_gpu_gps_kernel * FMA_SLOTS_PER_GATE / p_int, "peak_kind": p_int_src,
                                    "slots_per_gate": FMA_SLOTS_PER_GATE, "butterflies_per_gate": BUTTERFLIES_PER_GATE,
                                    "modmul_per_gate_single_modulus_basis": MODMUL_PER_GATE,
                                    "note": "slot weights measured on B200: IMAD 1, IMAD.HI 2, IMAD.WIDE 2.5 (profiles/intpipe_r01.json); "
                                            f"butterfly = 2 IMAD + 1 IMAD.HI = 4 slots; this design runs {fig['transforms_per_cmux']} transforms per CMUX"}
        if world == 1 and not args.no_cpu_baseline:
            try:
                nthreads = os.cpu_count() or 1
                gps1, secs1, B1, ok1, Kc = cpu_reference_run(1, 16)
                gpsN, secsN, BN, okN, _ = cpu_reference_run(nthreads, max(1, BATCH // nthreads), Kc)   # the whole 1024-gate batch
                line["cpu_baseline"] = {"value": gpsN, "unit": "gates/s", "cores": nthreads, "kind": "port",
                                        "value_1core": gps1, "ms_per_gate_1core": 1e3 / gps1, "decrypt_ok": bool(ok1 and okN),
                                        "sample": f"{BN} NAND gates on {nthreads} threads ({secsN:.1f} s) and {B1} gates on 1 thread "
                                                  f"({secs1:.1f} s); reference's own spqlios FFT (oracle/_ref) + C restatement of the "
                                                  f"Rust glue (no Rust toolchain in the image)"}
            except Exception as ex:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": "gates/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
        print(json.dumps(line), file=real_stdout, flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if wrong == 0 else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3 / 4 / 5 extra measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        sys.exit(run_reference(args))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29541", os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
               str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []) + (["--no-extras"] if args.no_extras else [])
        sys.exit(subprocess.call(cmd))
    sys.exit(run_gpu(args))


if __name__ == "__main__":
    main()
