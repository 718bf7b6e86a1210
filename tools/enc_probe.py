"""Dev tool: time the K7 encoding of a 16384^2 quarter tensor -- standalone (acetn_b200_i8_encode: column pass + row encoder) and as the
sweep produces it (acetn_b200_quarter_tensor_enc minus acetn_b200_quarter_tensor) --."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from acetn_b200 import ops
from acetn_b200.synthetic import random_ipeps

D, chi = 8, 256
ip = random_ipeps(1, 1, D, chi, 2, seed=0, device="cuda")
st = ip[(0, 0)]
qa = (st['C'][0], st['E'][0], st['E'][3], st['A'])
Q, _ = ops.quarter_tensor(*qa, normalize=False)
enc = ops.i8_encode(Q)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


t_alone = timed(lambda: ops.i8_encode(Q, storage=enc.storage))
t_q = timed(lambda: ops.quarter_tensor(*qa, normalize=False, out=Q.view(-1)))
t_qe = timed(lambda: ops.quarter_tensor(*qa, normalize=False, out=Q.view(-1), enc_storage=enc.storage))
print(f"i8_encode alone {t_alone:.3f} ms; quarter tensor {t_q:.3f} ms, with encoding {t_qe:.3f} ms "
      f"-> encoding inside the call {t_qe - t_q:.3f} ms", flush=True)
