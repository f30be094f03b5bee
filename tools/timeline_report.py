"""Summarises gpurun_out/r2_w_timeline_w*.txt.gz (tools/timeline_probe.py on a -DWF_TIMELINE build)."""
import gzip, sys, numpy as np
for path in sys.argv[1:]:
    lines = gzip.open(path, 'rt').read().split('\n')
    idx = [i for i, l in enumerate(lines) if l.startswith('FRAME')]
    seg = lines[idx[-2] + 1:idx[-1]]
    print(path, lines[idx[-1]])
    a = np.array([[int(x) for x in l.split()[1:]] for l in seg if l.startswith('TL')], dtype=np.int64)
    for m, name in ((0, 'primary'), (1, 'bounce'), (2, 'tail launch 1'), (3, 'tail launch 2')):
        b = a[(a[:, 0] == m) & (a[:, 2] > 0)]
        if len(b) == 0: continue
        t0 = b[:, 2].min()
        st = (b[:, 2] - t0) / 1e3; ex = (b[:, 3] - t0) / 1e3; en = (b[:, 4] - t0) / 1e3
        pc = lambda x, q: np.percentile(x, q)
        print(f" {name}: warps {len(b)}  launch length {en.max():.1f} us  rays {b[:, 5].sum()}  warps with work {(b[:, 5] > 0).sum()}  first start {(t0 - a[a[:, 2] > 0][:, 2].min()) / 1e3:.1f} us after the frame's first")
        print("   start   med/max           %.1f %.1f" % (np.median(st), st.max()))
        print("   exhaust min/p1/med/p99/max %.1f %.1f %.1f %.1f %.1f" % (ex.min(), pc(ex, 1), np.median(ex), pc(ex, 99), ex.max()))
        print("   exit    min/p1/p10/med/p90/p99/max %.1f %.1f %.1f %.1f %.1f %.1f %.1f" % (en.min(), pc(en, 1), pc(en, 10), np.median(en), pc(en, 90), pc(en, 99), en.max()))
        T = en.max()
        ts = np.linspace(0, T, 25)
        print("   warps alive at t (us): " + "  ".join("%.0f:%d" % (t, ((st <= t) & (en > t)).sum()) for t in ts))
