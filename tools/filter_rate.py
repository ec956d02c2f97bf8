"""Measured rates of the read_data filters on one B200, scipy on one host core beside them on a bounded
sample: python tools/filter_rate.py [N] -> one JSON line (kept under profiles/)."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import scipy.signal as signal
from system_identification_b200 import filters

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
CH = 48
rng = np.random.default_rng(9)
x = rng.normal(0, 1.0, (CH, N))
xd = torch.from_numpy(x).cuda()
b, a = filters.butter_lowpass(5, 0.15)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ms = []
    for it in range(7):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), out


ms_f, yf = timed(lambda: filters.filtfilt(b, a, xd))
ms_s, ys = timed(lambda: filters.savgol_filter(xd, 21, 5))
ns = min(N, 100_000)
t0 = time.perf_counter(); rf = signal.filtfilt(b, a, x[:, :ns], axis=1); t_f = time.perf_counter() - t0
t0 = time.perf_counter(); rs = signal.savgol_filter(x[:, :ns], 21, 5); t_s = time.perf_counter() - t0
# interior samples of the prefix are independent of where the signal ends to far below the tolerance
m = ns - 2000
assert np.abs(yf[:, :m].cpu().numpy() - rf[:, :m]).max() <= 1e-10 * np.abs(rf).max()
assert np.abs(ys[:, :m].cpu().numpy() - rs[:, :m]).max() <= 1e-10 * np.abs(rs).max()
# filtfilt: forward + backward, each reads the stream twice (kernels A and B) and writes it once: 6 x 8 B per sample
fb = 6 * 8 * CH * N
sb = 2 * 8 * CH * N
print(json.dumps({"workload": f"{CH} channels x {N} samples fp64, one B200",
                  "filtfilt_ms": ms_f, "filtfilt_algorithmic_GBps": fb / ms_f / 1e6, "filtfilt_samples_per_s": CH * N / ms_f * 1e3,
                  "savgol_ms": ms_s, "savgol_algorithmic_GBps": sb / ms_s / 1e6,
                  "scipy_filtfilt_samples_per_s": CH * ns / t_f, "scipy_savgol_samples_per_s": CH * ns / t_s,
                  "scipy_sample": f"{CH} x {ns}, one host core"}))
