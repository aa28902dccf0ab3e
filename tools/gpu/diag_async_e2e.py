import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import rustfhe_b200 as R
from rustfhe_b200 import _capi as K
seed = 0x5EED0001
sk = R.SecretKeys.generate(seed)
tfhe = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed)
eng = tfhe.engine
B, NROT = 1024, 8
bits = np.random.default_rng(1).integers(0, 2, B * NROT).astype(np.uint8)
cx = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, bits, seed=5, ct_index0=0)
hx = torch.from_numpy(cx.view(np.int32)).pin_memory(); hy = torch.from_numpy(cx.view(np.int32).copy()).pin_memory()
hx_np, hy_np = hx.numpy().view(np.uint32), hy.numpy().view(np.uint32)
houts = [torch.empty((B, 636), dtype=torch.int32).pin_memory() for _ in range(4)]
houts_np = [h.numpy().view(np.uint32) for h in houts]
lib = K.lib()
def step(it):
    o = (it % NROT) * B
    t0 = time.perf_counter()
    rc = lib.tfhe_b200_gate_batch_async(eng._ctx, K.NAND, K.ptr(hx_np[o:o + B]), K.ptr(hy_np[o:o + B]), K.ptr(houts_np[it % 4]), B)
    assert rc == 0
    return (time.perf_counter() - t0) * 1e3
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    hs = [step(it) for it in range(10)]
    t1 = time.perf_counter()
    eng.sync()
    t2 = time.perf_counter()
    print(f"async x10: enqueue {1e3*(t1-t0):.1f} ms, total {1e3*(t2-t0):.1f} ms; per-call host ms: {' '.join(f'{h:.2f}' for h in hs)}", flush=True)
eng.close()
