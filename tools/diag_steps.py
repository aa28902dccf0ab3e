import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rustfhe_b200 as R
from rustfhe_b200 import _capi as K
seed = 0x5EED0001
sk = R.SecretKeys.generate(seed)
tfhe = R.TFHE.new(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed)
eng = tfhe.engine
rng = np.random.default_rng(1)
B = 1024
bx = rng.integers(0, 2, B).astype(np.uint8); by = rng.integers(0, 2, B).astype(np.uint8)
cx = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, bx, seed=5, ct_index0=0)
cy = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, by, seed=6, ct_index0=0)
dev = torch.device("cuda", 0)
dx = torch.from_numpy(cx.view(np.int32)).to(dev); dy = torch.from_numpy(cy.view(np.int32)).to(dev)
do = [torch.empty_like(dx), torch.empty_like(dx)]
s0 = torch.cuda.current_stream(); s1 = torch.cuda.Stream(device=dev); s2 = torch.cuda.Stream(device=dev)
def run(streams, n, label):
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter(); host = []
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for s in streams[1:]: s.wait_event(e0)
    for it in range(n):
        st = streams[it % len(streams)]
        h0 = time.perf_counter()
        eng.gate_batch_device(K.NAND, dx.data_ptr(), dy.data_ptr(), do[it % 2].data_ptr(), B, st.cuda_stream)
        host.append((time.perf_counter() - h0) * 1e3)
    for s in streams[1:]:
        j = torch.cuda.Event(); j.record(s); streams[0].wait_event(j)
    e1.record(streams[0])
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{label}: {n} steps {ms:.2f} ms -> {n*B/ms*1e3:.0f} gates/s; host ms per call: {' '.join(f'{h:.2f}' for h in host)}", flush=True)
for rep in range(2):
    run([s0], 10, "default stream")
    run([s1], 10, "one side stream")
    run([s1, s2], 10, "two side streams")
    run([s0, s1], 10, "default + side")
