"""Diagnostic: bidirectional PCIe copy rates, one transfer vs chunked vs chunked with cross-stream dependencies (the
shape of the int16 host pipeline).  torch is used for pinned memory / streams only."""
import torch, time
MB = 1 << 20
n = 115_200_000
hin = torch.empty(n, dtype=torch.uint8).pin_memory(); hin.random_(0, 255)
hout = torch.empty(n, dtype=torch.uint8).pin_memory()
din = torch.empty(n, dtype=torch.uint8, device="cuda")
dout = torch.empty(n, dtype=torch.uint8, device="cuda")
sa, sb, sk = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

def run(chunk, dep, kern, reps=8):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream())
        sa.wait_event(e0); sb.wait_event(e0); sk.wait_event(e0)
        for o in range(0, n, chunk):
            c = min(chunk, n - o)
            with torch.cuda.stream(sa):
                din[o:o + c].copy_(hin[o:o + c], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(sa)
            if kern:
                with torch.cuda.stream(sk):
                    sk.wait_event(ev)
                    dout[o:o + c].copy_(din[o:o + c])          # a device-side pass over the chunk
                    ev = torch.cuda.Event(); ev.record(sk)
            with torch.cuda.stream(sb):
                if dep: sb.wait_event(ev)
                hout[o:o + c].copy_(dout[o:o + c], non_blocking=True)
        e1.record(sa); e2.record(sb)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
    return best

for chunk_mb in (110, 56, 28, 14, 7, 3.5):
    chunk = int(chunk_mb * MB)
    print(f"chunk {chunk_mb:6.1f} MB: independent {run(chunk, False, False):.3f} ms   d2h after h2d {run(chunk, True, False):.3f} ms   "
          f"+ device pass {run(chunk, True, True):.3f} ms", flush=True)
# one direction only, chunked
def one(chunk, h2d):
    best = 1e9
    for _ in range(6):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(sa):
            e0.record(sa)
            for o in range(0, n, chunk):
                c = min(chunk, n - o)
                if h2d: din[o:o + c].copy_(hin[o:o + c], non_blocking=True)
                else: hout[o:o + c].copy_(dout[o:o + c], non_blocking=True)
            e1.record(sa)
        torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for chunk_mb in (110, 14, 3.5):
    print(f"chunk {chunk_mb:6.1f} MB: h2d only {one(int(chunk_mb*MB), True):.3f} ms   d2h only {one(int(chunk_mb*MB), False):.3f} ms", flush=True)
