"""Host memory bandwidth of the box: multi-threaded memcpy (numpy releases the GIL), GB/s of payload."""
import numpy as np, threading, time, os
n = 64 << 20
for T in (1, 2, 4, 8, 16):
    src = [np.ones(n, np.uint8) for _ in range(T)]; dst = [np.empty(n, np.uint8) for _ in range(T)]
    for d, s in zip(dst, src): np.copyto(d, s)
    def work(i):
        for _ in range(8): np.copyto(dst[i], src[i])
    th = [threading.Thread(target=work, args=(i,)) for i in range(T)]
    t0 = time.perf_counter(); [t.start() for t in th]; [t.join() for t in th]; dt = time.perf_counter() - t0
    print("threads %2d: memcpy %.1f GB/s payload (%.1f GB/s read+write)" % (T, T * 8 * n / dt / 1e9, 2 * T * 8 * n / dt / 1e9), flush=True)
