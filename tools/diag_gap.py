import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rustfhe_b200 as R
from rustfhe_b200 import _capi as K
seed = 0x5EED0001
sk = R.SecretKeys.generate(seed)
tfhe = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed)
eng = tfhe.engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
bits = torch.randint(0, 2, (B,), dtype=torch.uint8, device=dev)
dx = torch.empty((B, 636), dtype=torch.int32, device=dev); dy = torch.empty_like(dx); do = torch.empty_like(dx)
st = torch.cuda.current_stream()
eng.encrypt_bits_device(1, 0, sk.s_key_tlwelv0, bits.data_ptr(), B, dx.data_ptr(), st.cuda_stream)
eng.encrypt_bits_device(1, B, sk.s_key_tlwelv0, bits.data_ptr(), B, dy.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
for rep in range(2):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(9)]
    eng.reset_stats()
    evs[0].record(st)
    for it in range(8):
        eng.gate_batch_device(K.NAND, dx.data_ptr(), dy.data_ptr(), do.data_ptr(), B, st.cuda_stream)
        evs[it + 1].record(st)
    torch.cuda.synchronize()
    s = eng.stats()
    print("per-call outer ms:", " ".join(f"{evs[i].elapsed_time(evs[i+1]):.2f}" for i in range(8)), "| avg br", round(s["avg_blind_rotate_ms"], 3), "ks", round(s["avg_keyswitch_ms"], 3), flush=True)
eng.close()
