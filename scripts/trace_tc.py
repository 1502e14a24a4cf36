"""Debug helper (GPU box): timeline of the fused tensor-path kernel from its in-kernel clock64 stamps.
AWB_TC_TRACE=1 python scripts/trace_tc.py  ->  per-stage cycle deltas of CTA 0 (epilogue thread 0, issuer)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SERIAL = "--serial" in sys.argv
if SERIAL:
    sys.argv.remove("--serial")
# diagnostic variants of the library: the production build carries no stamps; --serial additionally makes the issuer
# wait for (and stamp) every contraction group
import glob
import subprocess
import __graft_entry__ as entry
flag, name = ("-DAWB_TC_SERIAL", "libawb_serial.so") if SERIAL else ("-DAWB_TC_TRACE_BUILD", "libawb_trace.so")
lib = os.path.join(entry.CSRC, name)
srcs = sorted(glob.glob(os.path.join(entry.CSRC, "*.cu"))) + glob.glob(os.path.join(entry.CSRC, "*.cuh"))
if not os.path.exists(lib) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in srcs):
    subprocess.check_call(["nvcc"] + entry.NVCC_FLAGS + [flag, "-o", lib] + sorted(glob.glob(os.path.join(entry.CSRC, "*.cu"))))
os.environ["AWB_LIB_PATH"] = lib
if SERIAL:
    os.environ["AWB_TC_TRACE"] = "2"
os.environ.setdefault("AWB_TC_TRACE", "1")
import torch
import awesome_b200 as A
from awesome_b200 import _lib

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W = 480, 640
torch.manual_seed(0)
m = A.ConvexNextNet(n_hidden_layers=L, precision="f16").cuda()
yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, W), indexing="ij")
un = torch.sigmoid((torch.sqrt(((xx - 0.52) / 0.27) ** 2 + ((yy - 0.47) / 0.31) ** 2) - 1) / 0.08).cuda()
grid = A.GridSpecHost("linspace", 1, H, W)
f = m.make_fitter(grid, un, A.LossConfig("mse"), A.OptimConfig("adam", lr=1e-3), use_graph=False)
f.run(5)
torch.cuda.synchronize()
lib = _lib.load()
n_cta = 4
buf = (C.c_ulonglong * (256 * n_cta))()
got = lib.awb_debug_tc_trace_read(buf, n_cta)
print("ctas read:", got)
for c in range(got):
    ep = [buf[c * 256 + i] for i in range(128)]
    ms = [buf[c * 256 + 128 + i] for i in range(128)]
    k0, k1, k2, k3, k4 = ep[120:125]
    ns = ep[126] - ep[125]
    print(f"--- CTA {c}: {k4 - k0} cycles in {ns} ns -> SM clock {(k4 - k0) / max(ns, 1) * 1e3:.0f} MHz")
    print(f"--- CTA {c}: setup {k1 - k0} | first stage + weights {k2 - k1} | tile loop {k3 - k2} | write-out {k4 - k3} cycles")
    ep = [x for x in ep[:120] if x]
    ms = [x for x in ms if x]
    t0 = min(ep[0], ms[0])
    print(f"--- CTA {c}: epilogue stamps (wait-done, arrive-done alternating), relative cycles")
    print(" ".join(str(x - t0) for x in ep[:40]))
    print("    deltas:", " ".join(str(b - a) for a, b in zip(ep[:40], ep[1:41])))
    print(f"--- CTA {c}: issuer stamps (first after prologue commit; then per stage: wake, issued)")
    print(" ".join(str(x - t0) for x in ms[:40]))
    print("    deltas:", " ".join(str(b - a) for a, b in zip(ms[:40], ms[1:41])))
    if SERIAL and L == 2:
        # per tile: wake, fwd1, issued | wake, fwd2, issued | wake, dgrad2, GO, wgrad2, PB2, issued | wake, dgrad1, wgrad1,
        # PB1, issued | wake, input', GIN, issued      (each duration includes one commit -> mbarrier -> wake round trip)
        names = ["wake", "fwd1", "-", "wake", "fwd2", "-", "wake", "dgrad2", "GO", "wgrad2", "PB2", "-", "wake", "dgrad1",
                 "wgrad1", "PB1", "-", "wake", "input'", "GIN", "-"]
        base = 1 + len(names)       # second tile
        d = [ms[base + i] - ms[base + i - 1] for i in range(len(names))]
        print("    serialised issue, second tile: " + "  ".join(f"{n} {x}" for n, x in zip(names, d) if n != "-"))
    nst = 2 * L + 1
    if len(ep) > 2 * nst * 3:
        per_tile = (ep[2 * nst * 3] - ep[2 * nst * 1]) / 2.0
        print(f"    cycles per tile (tiles 1..2): {per_tile:.0f}")
