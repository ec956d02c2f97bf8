"""Measured rates of the ingest row on one B200, with the reference's CPU calls timed beside them on a
bounded sample: python tools/ingest_rate.py [N]  -> one JSON line (kept under profiles/).

  parse    sysid_dat_scan + sysid_dat_parse on the '%.6f' text of an (18 x N) array, text resident in HBM,
           CUDA events; roofline = HBM: bytes read twice (count pass + parse pass) + 8 B written per field
  loadtxt  np.loadtxt(dtype=float32) of the first 18 x 20 000 fields of the same text (the reference's read_data call)
  fd       sysid_fd_rate on 15 channels x N against the reference's df.at row loop restated (oracle) on 2 000 rows
"""
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C

from oracle import ingest as oi
from system_identification_b200 import _lib, ingest
from system_identification_b200.ops import _ptr, _stream


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rng = np.random.default_rng(5)
    x = rng.normal(0, 3.0, (18, N))
    t0 = time.perf_counter()
    buf = io.BytesIO()
    np.savetxt(buf, x, delimiter="\t", fmt="%.6f")
    text = buf.getvalue()
    t_save = time.perf_counter() - t0
    lib = _lib.load()
    n = len(text)
    dtext = torch.empty(n + 16, dtype=torch.uint8, device="cuda")
    dtext[:n].copy_(torch.frombuffer(bytearray(text), dtype=torch.uint8))
    ws = torch.empty(lib.sysid_dat_workspace_bytes(n), dtype=torch.uint8, device="cuda")
    out = torch.empty((18, N), dtype=torch.float64, device="cuda")
    dims = (C.c_int64 * 2)()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms_scan, ms_parse = [], []
    for it in range(6):
        flush.fill_(it)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.check(lib.sysid_dat_scan(_ptr(dtext), n, 9, _ptr(ws), ws.numel(), dims, _stream()))
        e[1].record()
        _lib.check(lib.sysid_dat_parse(_ptr(dtext), n, 9, _ptr(ws), ws.numel(), 18, N, _ptr(out), N, 1, None, _stream()))
        e[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            ms_scan.append(e[0].elapsed_time(e[1])); ms_parse.append(e[1].elapsed_time(e[2]))
    assert (int(dims[0]), int(dims[1])) == (18, N)
    ms = float(np.mean(ms_scan) + np.mean(ms_parse))
    alg_bytes = 2 * n + 8 * 18 * N
    # end to end from the host bytes (pageable -> pinned -> device -> parse), wall clock
    t0 = time.perf_counter()
    got = ingest.load_dat(text)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    # CPU reference on a bounded sample
    ns = min(N, 20000)
    sbuf = io.BytesIO()
    np.savetxt(sbuf, x[:, :ns], delimiter="\t", fmt="%.6f")
    t0 = time.perf_counter()
    ref = np.loadtxt(io.BytesIO(sbuf.getvalue()), delimiter="\t", dtype=np.float32)
    t_loadtxt = time.perf_counter() - t0
    assert np.array_equal(got[:, :ns].cpu().numpy(), ref.astype(np.float64))
    # finite differences
    tick = np.cumsum(rng.integers(1, 4, N)).astype(np.float64)
    dx = torch.from_numpy(x[:15]).cuda()
    dt = torch.from_numpy(tick).cuda()
    y = torch.empty_like(dx)
    ms_fd = []
    for it in range(6):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.sysid_fd_rate(_ptr(dt), _ptr(dx), _ptr(y), 15, N, N, N, 1000.0, _stream()))
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ms_fd.append(e0.elapsed_time(e1))
    nf = 2000
    t0 = time.perf_counter()
    yo = oi.fd_rate(tick[:nf], x[:15, :nf], 1000.0)
    t_fd_cpu = time.perf_counter() - t0
    assert np.array_equal(y[:, 1:nf].cpu().numpy(), yo[:, 1:])
    fd_bytes = 8 * (2 * 15 * N + N)
    print(json.dumps({
        "workload": f"18 x {N} '%.6f' tab-separated text ({n / 1e6:.1f} MB), one B200",
        "parse_ms": ms, "scan_ms": float(np.mean(ms_scan)), "convert_ms": float(np.mean(ms_parse)),
        "parse_text_GBps": n / ms / 1e6, "parse_fields_per_s": 18 * N / ms * 1e3,
        "parse_roofline": {"bound": "hbm", "algorithmic_bytes": alg_bytes, "achieved_GBps": alg_bytes / ms / 1e6},
        "load_dat_e2e_s_from_host_bytes": t_e2e,
        "np_loadtxt_fields_per_s": 18 * ns / t_loadtxt, "np_loadtxt_sample": f"18 x {ns}", "np_savetxt_s_full": t_save,
        "fd_ms": float(np.mean(ms_fd)), "fd_GBps": fd_bytes / float(np.mean(ms_fd)) / 1e6,
        "fd_reference_loop_rows_per_s": nf / t_fd_cpu, "fd_reference_sample": f"15 channels x {nf} rows (oracle restatement of the df.at loop)",
    }))


if __name__ == "__main__":
    main()
