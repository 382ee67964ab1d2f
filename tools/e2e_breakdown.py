"""Where does the streamed host path spend its time?  Same text, same box, one process:
   pure copies (H2D of the text in slabs, D2H of the matrix in pitched slabs), then hb_parse_stream_host with and
   without its outputs.  python tools/e2e_breakdown.py [variants] [samples]"""
import os, sys, time, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from haplohyped_varawareml_b200 import capi

V = int(sys.argv[1]) if len(sys.argv) > 1 else 1_100_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2504
spec = capi.synth_spec(V, S, seed=42)
T = int(capi.lib().hb_synth_body_bytes(spec))
text = torch.empty(T + 256, dtype=torch.uint8, device="cuda")
capi.check(capi.lib().hb_synth_device(spec, text.data_ptr(), T, 0, None))
host = torch.empty(T, dtype=torch.uint8).pin_memory()
host.copy_(text[:T]); torch.cuda.synchronize()
out0 = torch.empty((S, V), dtype=torch.int8).pin_memory()
out1 = torch.empty((S, V), dtype=torch.int8).pin_memory()
sites = [torch.empty(V, dtype=torch.int32).pin_memory(), torch.empty(V, dtype=torch.int32).pin_memory(),
         torch.empty(V, dtype=torch.uint8).pin_memory(), torch.empty(V, dtype=torch.uint8).pin_memory()]
res = {}
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    res[name] = [round(x, 1) for x in ts]
slab = 256 << 20
def h2d_only():
    for off in range(0, T, slab):
        n = min(slab, T - off)
        text[off:off + n].copy_(host[off:off + n], non_blocking=True)
timed("h2d_text_slabs_ms", h2d_only)
timed("h2d_text_one_copy_ms", lambda: text[:T].copy_(host, non_blocking=True))
opts = capi.Parse._opts(S, "chr22", False, True, 0, 0, None)
nrec = C.c_uint64()
def stream(outs, slab_bytes=slab, want_gt=True):
    o = capi.Parse._opts(S, "chr22", False, want_gt, 0, 0, None)
    a = [out0.data_ptr(), out1.data_ptr()] if outs else [None, None]
    s = [x.data_ptr() for x in sites] if outs else [None] * 4
    capi.check(capi.lib().hb_parse_stream_host(host.data_ptr(), T, C.byref(o), slab_bytes, a[0], a[1], V, s[0], s[1], s[2], s[3],
                                               None, None, C.byref(nrec), None))
timed("stream_full_ms", lambda: stream(True))
timed("stream_no_outputs_ms", lambda: stream(False))
timed("stream_no_gt_ms", lambda: stream(False, want_gt=False))
timed("stream_full_1GiB_ms", lambda: stream(True, 1 << 30))
timed("stream_full_64MiB_ms", lambda: stream(True, 64 << 20))
# BGZF streamed
hdr = capi.synth_header(spec)
full = np.empty(len(hdr) + T, np.uint8); full[:len(hdr)] = np.frombuffer(hdr, np.uint8); full[len(hdr):] = host.numpy()
bg = capi.bgzf_compress_host(full, 6); del full
bgp = torch.empty(bg.size, dtype=torch.uint8).pin_memory(); bgp.numpy()[:] = bg; del bg
def bstream(slab_bytes, outs=True):
    a = [out0.data_ptr(), out1.data_ptr()] if outs else [None, None]
    s = [x.data_ptr() for x in sites] if outs else [None] * 4
    capi.check(capi.lib().hb_parse_stream_bgzf_host(bgp.data_ptr(), bgp.numel(), b"chr22", 1, 0, slab_bytes, a[0], a[1], V, s[0], s[1], s[2], s[3],
                                                    None, None, C.byref(nrec), None))
timed("bgzf_stream_256MiB_ms", lambda: bstream(256 << 20))
timed("bgzf_stream_1GiB_ms", lambda: bstream(1 << 30))
timed("bgzf_stream_1GiB_no_outputs_ms", lambda: bstream(1 << 30, False))
# D2H only: the matrix of a resident parse
p = capi.Parse.from_device(text.data_ptr(), T, S, region="chr22")
timed("fetch_matrix_ms", lambda: capi.check(capi.lib().hb_parse_fetch_matrix(p._h, out0.data_ptr(), out1.data_ptr())))
print(json.dumps(res))
