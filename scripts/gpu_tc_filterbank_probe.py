"""Measured tensor-core row for the mel filterbank: what the DENSE [T x 257] x [257 x nMels] product costs on
the B200 tensor cores, with the operand precision the 1e-4 log-mel bar needs.

The product is run by cuBLAS (tcgen05 kernels on sm_100) on a power tile the size of BASELINE's hour, as
  (a) one bf16 GEMM                          (accuracy only: bf16 operands)
  (b) bf16 x 3 split, six GEMMs (hi*hi, hi*mid, mid*hi, hi*lo, lo*hi, mid*mid), fp32 accumulate
  (c) one TF32 GEMM and 3xTF32 (three GEMMs)
and compared with the float32 banded product the kernel uses (numerics: float64 reference of the same contraction).
Times are for the GEMMs ALONE with operands already split and resident in HBM — a lower bound for any fused tcgen05 stage,
which would additionally have to split and re-lay-out the power tile in shared memory.  Not part of the product or of bench.py.
"""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import oracle as O

T, BINS, NM = 360001, 257, 80
KP = 272
dev = torch.device("cuda")
torch.manual_seed(0)
# a realistic power tile: the oracle's own spectrum statistics (log-normal, 60 dB of range)
power = torch.exp(torch.randn(T, KP, device=dev) * 3.0 - 4.0)
power[:, BINS:] = 0
fb = torch.zeros(KP, NM, device=dev)
fb[:BINS] = torch.from_numpy(O.mel_filterbank(512, NM).T.copy()).to(dev)
ref = torch.log((power.double() @ fb.double()) + 2.0 ** -24)

def split3(x):
    hi = x.to(torch.bfloat16); r = x - hi.float()
    mid = r.to(torch.bfloat16); lo = (r - mid.float()).to(torch.bfloat16)
    return hi, mid, lo

def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

rows = []
f32 = torch.log(power @ fb + 2.0 ** -24)
torch.backends.cuda.matmul.allow_tf32 = False
rows.append(("fp32 SIMT GEMM (dense)", timeit(lambda: power @ fb), (torch.log(power @ fb + 2.0 ** -24) - ref).abs().max().item()))
torch.backends.cuda.matmul.allow_tf32 = True
rows.append(("TF32 tensor GEMM x1", timeit(lambda: power @ fb), (torch.log(power @ fb + 2.0 ** -24) - ref).abs().max().item()))
torch.backends.cuda.matmul.allow_tf32 = False
ph, pm, pl = split3(power); fh, fm, fl = split3(fb)
# torch returns bf16 from a bf16 matmul (the fp32 accumulator is rounded on the way out): TIMES come from the real bf16
# GEMMs, ACCURACY from float32 GEMMs of the same bf16-rounded operands (products exact, fp32 accumulate — what a tcgen05
# kind::f16 MMA with an fp32 accumulator in TMEM delivers)
one = lambda: (ph @ fh).float()
one_acc = lambda: ph.float() @ fh.float()
rows.append(("bf16 tensor GEMM x1", timeit(one), (torch.log(one_acc() + 2.0 ** -24) - ref).abs().max().item()))
def six():
    acc = (ph @ fh).float(); acc += (ph @ fm).float(); acc += (pm @ fh).float()
    acc += (ph @ fl).float(); acc += (pl @ fh).float(); acc += (pm @ fm).float()
    return acc
def six_acc():
    P = [x.float() for x in (ph, pm, pl)]; F = [x.float() for x in (fh, fm, fl)]
    return P[0] @ F[0] + P[0] @ F[1] + P[1] @ F[0] + P[0] @ F[2] + P[2] @ F[0] + P[1] @ F[1]
rows.append(("bf16x3 tensor GEMM x6 (operands pre-split)", timeit(six), (torch.log(six_acc() + 2.0 ** -24) - ref).abs().max().item()))
rows.append(("  + splitting the power tile into 3 bf16 planes (elementwise)", timeit(lambda: split3(power)), float("nan")))
print(f"dense filterbank product on tensor cores, T={T}, K={KP} (257 bins), N={NM}; ms per audio-hour, max |d log-mel| vs float64")
for name, ms, err in rows:
    print(f"  {name:62s} {ms:8.3f} ms   {err:.2e}")
print("for comparison: the whole fused float32-pair mel kernel is 0.337 ms per audio-hour; its banded float32 filterbank + log stage is "
      "~24 % of that (0.08 ms), max |d| 2e-6")
