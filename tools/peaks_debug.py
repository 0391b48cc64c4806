import sys, warnings; sys.path.insert(0, '/root/repo')
import numpy as np
from scipy.signal import find_peaks
warnings.simplefilter("ignore")
from adapted_b200.detect import find_peaks_device
import tests.test_gpu_peaks_adversarial as T
from oracle import detect_ref
x = np.zeros(60); x[[10, 14, 30, 37, 50]] = [5, 5, 3, 3, 4]
print("ties:", find_peaks_device([x], mode=0, distance=10, prominence=0.0, width=0.0, want=32)[0], find_peaks(x, distance=10, prominence=0.0, width=0.0)[0])
print("ties nodist:", find_peaks_device([x], mode=0, distance=0, prominence=0.0, width=0.0, want=32)[0], find_peaks(x, prominence=0.0, width=0.0)[0])
x2 = np.zeros(60); x2[[10, 14, 30, 37, 50]] = [5, 4, 3, 2, 4]
print("no ties:", find_peaks_device([x2], mode=0, distance=10, prominence=0.0, width=0.0, want=32)[0], find_peaks(x2, distance=10, prominence=0.0, width=0.0)[0])
for kind in ("plateaus", "special"):
    shown = 0
    for seed in range(60):
        rng = np.random.default_rng(seed)
        traces = T._traces(rng, 24, kind)
        got = find_peaks_device(traces, mode=0, distance=10, prominence=1.0, width=10.0, rel_height=0.5, want=32, nan_to_num=True)
        for x, g in zip(traces, got):
            y = np.nan_to_num(x, nan=0)
            pk, _ = find_peaks(y, distance=10, prominence=1.0, width=10, rel_height=0.5)
            allmax, _ = find_peaks(y)
            if allmax.size > 16 and ((allmax[1:] - allmax[:-1] < 10) & (y[allmax][1:] == y[allmax][:-1])).any():
                continue
            if not np.array_equal(g, pk[:32]) and shown < 3:
                shown += 1
                d = sorted(set(g.tolist()) ^ set(pk[:32].tolist()))
                print(kind, seed, "n", x.size, "nmax", allmax.size, "diff", d[:6])
                for p in d[:2]:
                    lo, hi = max(p - 12, 0), min(p + 13, x.size)
                    print("   around", p, "maxima", [int(m) for m in allmax if lo <= m < hi], "vals", np.round(y[lo:hi], 3).tolist())
