"""The pipelined rebuild decision of mc_step (engine.cu) replayed on an oracle trajectory: kick_drift raises the flag when
an atom's displacement since the last build exceeds skin/2 minus 1.5 drifts of look-ahead, and the host acts on the flag
of the PREVIOUS step.  Claim checked here, step by step: the list in force at a force evaluation was never built more
than skin/2 away from any atom's current position -- and the look-ahead costs only a few extra rebuilds."""
import numpy as np

from molchanica_b200 import workloads as W


def test_stale_list_is_never_used_beyond_half_the_skin(oracle):
    w = W.lj_fluid(m=9, temp_k=600.0)            # hot: the criterion fires every ~10 steps
    dt, skin = w["dt"], w["skin"]
    ext = np.asarray(w["box_ext"], np.float64)
    x, v = w["xyzq"].copy(), w["vel"].copy()
    xs = [x[:, :3].astype(np.float64)]
    for _ in range(160):
        r = oracle.md_run(w, 1, precision=64, xyzq=x, vel=v)
        x, v = r["xyzq"], r["vel"]
        xs.append(x[:, :3].astype(np.float64))

    def disp(a, b):
        d = a - b
        d -= np.rint(d / ext) * ext
        return np.sqrt((d * d).sum(1))

    def replay(lookahead, late):
        """late = True: the host sees the flag one step late (the pipelined path); returns (rebuilds, worst displacement / (skin/2))."""
        xref, rebuilds, worst = xs[0], 0, 0.0
        prev_flag, skip_prev = None, False
        for s in range(1, len(xs)):
            speed = disp(xs[s], xs[s - 1]) / dt                      # the half-step velocity kick_drift drifts with
            thr = 0.5 * skin - lookahead * speed * dt
            flag = bool(np.any((thr <= 0) | (disp(xs[s], xref) > thr)))
            if late:
                rebuild = prev_flag is not None and not skip_prev and prev_flag
                prev_flag, skip_prev = flag, rebuild
            else:
                rebuild = flag
            if rebuild:
                xref = xs[s]
                rebuilds += 1
            else:
                worst = max(worst, float(disp(xs[s], xref).max()) / (0.5 * skin))
        return rebuilds, worst

    ideal, w_ideal = replay(0.0, late=False)         # synchronous flag, no look-ahead: the fewest rebuilds that are safe
    piped, w_piped = replay(1.5, late=True)          # what mc_step does
    naive, w_naive = replay(0.0, late=True)          # acting late WITHOUT look-ahead is not safe
    assert w_ideal <= 1.0 and w_piped <= 1.0, (w_ideal, w_piped)
    assert w_naive > 1.0
    assert ideal >= 8 and piped <= 1.35 * ideal + 1, (ideal, piped)


def test_adaptive_interval_of_decomposed_runs(oracle):
    """comm_rebuild's schedule rule (comm.cu): after an interval of k steps whose largest displacement was d, the next
    interval is floor(0.75 k (skin/2) / d), at most doubling (+ 1) per build, never below 4.  Replayed on an oracle
    trajectory it climbs to the useful range and never overshoots skin/2 -- whereas a target of 85 % (the first choice,
    which the 1M-atom runs happened to survive) does overshoot on this small hot system."""
    w = W.lj_fluid(m=9, temp_k=300.0)
    dt, skin = w["dt"], w["skin"]
    ext = np.asarray(w["box_ext"], np.float64)
    x, v = w["xyzq"].copy(), w["vel"].copy()
    xs = [x[:, :3].astype(np.float64)]
    for _ in range(400):
        r = oracle.md_run(w, 1, precision=64, xyzq=x, vel=v)
        x, v = r["xyzq"], r["vel"]
        xs.append(x[:, :3].astype(np.float64))

    def disp_max(a, b):
        d = a - b
        d -= np.rint(d / ext) * ext
        return float(np.sqrt((d * d).sum(1)).max())

    def replay(target):
        interval, s, xref, k = 10, 0, xs[0], 0
        intervals, fracs = [], []
        while s + 1 < len(xs):
            s += 1
            k += 1
            if k >= interval:
                frac = disp_max(xs[s], xref) / (0.5 * skin)
                fracs.append(frac)
                intervals.append(interval)
                want = target * k / frac if frac > 1e-6 else 2.0 * k + 1
                interval = int(max(4.0, min(200.0, np.floor(min(want, 2.0 * k + 1.0)))))
                xref, k = xs[s], 0
        return intervals, fracs

    intervals, fracs = replay(0.75)
    assert len(intervals) >= 6
    assert max(fracs) <= 0.95, fracs                      # no interval came close to overshooting skin/2
    assert intervals[-1] > intervals[0] and 0.55 < fracs[-1] <= 0.95   # it has climbed to the useful range
    assert max(replay(0.85)[1]) > 0.97                    # the earlier target leaves no margin here (0.9998 of skin/2)
