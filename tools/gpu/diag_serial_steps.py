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
B, NROT = 1024, 32
dev = torch.device("cuda", 0)
bits = torch.randint(0, 2, (B * NROT,), dtype=torch.uint8, device=dev)
dx = torch.empty((B * NROT, 636), dtype=torch.int32, device=dev); dy = torch.empty_like(dx); do = torch.empty((B, 636), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
eng.encrypt_bits_device(1, 0, sk.s_key_tlwelv0, bits.data_ptr(), B * NROT, dx.data_ptr(), st.cuda_stream)
eng.encrypt_bits_device(1, B * NROT, sk.s_key_tlwelv0, bits.data_ptr(), B * NROT, dy.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
def step(it):
    o = (it % NROT) * B
    eng.gate_batch_device(K.NAND, dx[o:o + B].data_ptr(), dy[o:o + B].data_ptr(), do.data_ptr(), B, st.cuda_stream)
for it in range(3): step(it)
torch.cuda.synchronize()
x = do.cpu().numpy()
for rep in range(2):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    eng.reset_stats()
    torch.cuda.synchronize()
    evs[0].record(st)
    for it in range(10):
        step(3 + it)
        evs[it + 1].record(st)
    torch.cuda.synchronize()
    s = eng.stats()
    print("per-call outer ms:", " ".join(f"{evs[i].elapsed_time(evs[i+1]):.2f}" for i in range(10)), "| total", round(evs[0].elapsed_time(evs[10]), 2), "avg br", round(s["avg_blind_rotate_ms"], 3), "ks", round(s["avg_keyswitch_ms"], 3), flush=True)
eng.close()
