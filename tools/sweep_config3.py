#!/usr/bin/env python3
"""BASELINE.json config 3, device resident: negacyclic poly-mul and external-product micro-benchmarks at N = 1024, batch sweep
1 .. 65536, against the integer-pipe and HBM rooflines.  GPU box only.

  work per product (this design): poly-mul = 7 transforms (1 digit + 3 key-slice forward, 3 inverse) + 3 x 1024 pointwise
  MACs; external product = 12 transforms + 36 x 1024 wide MACs + 6144 reductions (= one CMUX step);
  the 12 key transforms of a TRGSW are paid once per CALL (shared TRGSW) or once per item (per-item TRGSW).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

BFLY = 5120                      # butterflies per 1024-point transform
SLOTS_XP = 12 * BFLY * 4 + 36864 * 2.5 + 6144 * 3          # FMA-heavy issue slots of one external product
SLOTS_PM = 7 * BFLY * 4 + 3 * 1024 * (2.5 + 3)            # of one exact negacyclic product


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-log2", type=int, default=16)
    args = ap.parse_args()
    import torch
    import rustfhe_b200 as R
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    try:
        p_int = json.load(open(os.path.join(ROOT, "profiles", "intpipe_r01.json")))["imad_lo"]["Gops_per_s"] * 1e9
    except Exception:
        p_int = 148 * 64 * 1.965e9
    eng = R.DeviceEngine(0)
    fft64 = eng.stats()["key_slices"] == 1
    # FFT64 arithmetic (the default): both kernels are bound by issue slots like the gate kernel (bench.py fft64_figures; a DFMA holds
    # the issue port for two cycles): cycles per product = 2 x FP64 instructions + the others
    issue_peak = 148 * 4 * 1.965e9
    issue_pm = 2 * 1468 + 660      # ncu, profiles/r02_ncu_polymul_f64.txt: FP64 + other warp instructions per product
    issue_xp = 2 * 4415 + 2620     # ncu, profiles/r02_ncu_external_product_f64.txt
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream()
    g = torch.Generator(device=dev).manual_seed(7)

    def u32(*shape):
        return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=dev, generator=g).to(torch.int32)

    def timed(fn, reps=5):
        for _ in range(5):          # warm every workspace slot of the ring
            fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best

    res = {"gpu": torch.cuda.get_device_name(0), "arithmetic": "FFT64" if fft64 else "NTT", "issue_peak_cycles_per_s": issue_peak, "int_peak_slots_per_s": p_int, "hbm_gbs": peaks["hbm_gbs"], "negacyclic_mul": [],
           "external_product_shared_trgsw": [], "external_product_per_item_trgsw": []}
    for lb in range(0, args.max_log2 + 1, 2):
        B = 1 << lb
        a = u32(B, 1024)
        d = torch.randint(-32, 32, (B, 1024), dtype=torch.int32, device=dev, generator=g)
        o = torch.empty_like(a)
        t = timed(lambda: eng.negacyclic_mul_batch_device(a.data_ptr(), d.data_ptr(), o.data_ptr(), B, st.cuda_stream))
        res["negacyclic_mul"].append({"batch": B, "seconds": t, "products_per_s": B / t,
                                      **({"issue_roofline_frac": B / t * issue_pm / issue_peak} if fft64 else {"int_roofline_frac": B / t * SLOTS_PM / p_int}),
                                      "hbm_roofline_frac": B / t * 3 * 4096 / (peaks["hbm_gbs"] * 1e9)})
        trl = u32(B, 2, 1024)
        out = torch.empty_like(trl)
        trg1 = u32(1, 6, 2, 1024)
        t = timed(lambda: eng.external_product_batch_device(trg1.data_ptr(), 1, trl.data_ptr(), out.data_ptr(), B, st.cuda_stream))
        res["external_product_shared_trgsw"].append({"batch": B, "seconds": t, "products_per_s": B / t,
                                                     **({"issue_roofline_frac": B / t * issue_xp / issue_peak} if fft64 else {"int_roofline_frac": B / t * SLOTS_XP / p_int}),
                                                     "hbm_roofline_frac": B / t * 2 * 8192 / (peaks["hbm_gbs"] * 1e9)})
        if B <= 4096:
            trgB = u32(B, 6, 2, 1024)
            t = timed(lambda: eng.external_product_batch_device(trgB.data_ptr(), B, trl.data_ptr(), out.data_ptr(), B, st.cuda_stream))
            res["external_product_per_item_trgsw"].append({"batch": B, "seconds": t, "products_per_s": B / t,
                                                           "note": "includes the 36 key-slice transforms of every item's TRGSW"})
        del a, d, o, trl, out
    eng.close()
    txt = json.dumps(res, indent=1)
    if args.out:
        open(args.out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main()
