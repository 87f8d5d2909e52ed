"""Per-source-line view of an ncu report: instructions, stall samples and shared-memory wavefronts per frame.
Usage: python scripts/ncu_lines.py gpurun_out/prof_mel.ncu-rep [units_per_launch] [top]"""
import subprocess, csv, sys
rep = sys.argv[1]; units = float(sys.argv[2]) if len(sys.argv) > 2 else 360001.0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; per = {}; hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
        wf = hdr.index("L1 Wavefronts Shared"); ex = hdr.index("L1 Wavefronts Shared Excessive"); continue
    if r[0] == "Function Name": continue
    if r[0] != "" and hdr:
        try: per[(cur, int(r[0]))] = (r[1].strip()[:84], float(r[ie] or 0), float(r[sm] or 0), float(r[wf] or 0), float(r[ex] or 0))
        except Exception: pass
ti = sum(v[1] for v in per.values()); ts = sum(v[2] for v in per.values()); tw = sum(v[3] for v in per.values()); te = sum(v[4] for v in per.values())
print(f"per unit: instr {ti/units:.1f}  smem wavefronts {tw/units:.1f} (excessive {te/units:.1f})  samples {ts:.0f}")
for (f, l), (src, n, s, w, e) in sorted(per.items(), key=lambda kv: -kv[1][2])[:top]:
    print(f"{f}:{l:4d} instr {n/units:6.1f} samp {100*s/ts:5.1f}% wf {w/units:6.1f} (+{e/units:4.1f})  {src}")
