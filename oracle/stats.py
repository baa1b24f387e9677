"""CPU oracle for the per-genotype controller and statistics -- TEST INFRASTRUCTURE ONLY.

numpy restatement of the reference's repeat-until-pass loop and its reductions
(monte_carlo(), /root/reference/sonics.pyx:47-269), with the third-party
arithmetic it calls restated from the published formulas:

  * scipy.stats.mannwhitneyu(x, y, alternative="greater") (scipy 1.5.2 pinned at
    conda_env.yml:13; asymptotic method): midranks over the concatenation,
    U = R_x - n1(n1+1)/2, tie-corrected sigma, continuity 0.5, p = sf(z).
    Call site sonics.pyx:168-172.
  * pandas DataFrame.quantile(q) (pandas 1.1.3, conda_env.yml:10): linear
    interpolation at (n-1)q over the sorted group, sentinels included.
    Call sites sonics.pyx:120,158,159,236,237.

Pinned against the reference itself: tests/golden/monte_carlo_rows.json holds
rows produced by the reference's own monte_carlo() (oracle/_ref) under
np.random.seed(s); tests/test_oracle_stats.py replays the same seeds through
oracle.run() + this module and requires the same rows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
import math

import numpy as np

SENTINEL = -999999.0


def round_schedule(max_n_reps, monte_carlo=False):
    """Cumulative checkpoints of sonics.pyx:79-82,193-199: 100, 500, 1000, 5000, ..."""
    out = []
    run = 0
    reps = max_n_reps if monte_carlo else (100 if max_n_reps > 100 else max_n_reps)
    rnd = 1
    while run < max_n_reps:
        run += reps
        out.append(run)
        reps = 4 * run if rnd % 2 == 1 else run
        if reps + run > max_n_reps:
            reps = max_n_reps - run
        rnd += 1
    return out


def quantile_linear(sorted_vals, q):
    """pandas/numpy 'linear' quantile on an ascending array."""
    n = len(sorted_vals)
    h = (n - 1) * q
    lo = int(math.floor(h))
    hi = min(lo + 1, n - 1)
    t = h - lo
    a, b = float(sorted_vals[lo]), float(sorted_vals[hi])
    # numpy's _lerp: a + (b-a)t, switched to b - (b-a)(1-t) for t >= 0.5
    d = b - a
    r = a + d * t
    if t >= 0.5:
        r = b - d * (1 - t)
    if d == 0 or t == 0:
        r = a
    return r


def mwu_greater(x, y):
    """One-sided asymptotic Mann-Whitney U test, H1: x stochastically greater than y.

    Returns (U, p).  Raises ValueError when every value is identical (scipy 1.5.2
    behaviour, swallowed by the reference at sonics.pyx:175-176)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    n1, n2 = len(x), len(y)
    allv = np.concatenate([x, y])
    order = np.argsort(allv, kind="mergesort")
    sv = allv[order]
    n = n1 + n2
    ranks = np.empty(n, dtype=np.float64)
    # midranks
    i = 0
    tie_sum = 0.0
    while i < n:
        j = i
        while j + 1 < n and sv[j + 1] == sv[i]:
            j += 1
        r = 0.5 * ((i + 1) + (j + 1))
        ranks[order[i:j + 1]] = r
        t = j - i + 1
        tie_sum += float(t) ** 3 - t
        i = j + 1
    R1 = ranks[:n1].sum()
    U1 = R1 - n1 * (n1 + 1) / 2.0
    if tie_sum == float(n) ** 3 - n:
        raise ValueError("All numbers are identical in mannwhitneyu")
    mu = n1 * n2 / 2.0
    s = math.sqrt(n1 * n2 / 12.0 * ((n + 1) - tie_sum / (n * (n - 1))))
    z = (U1 - mu - 0.5) / s
    p = 0.5 * math.erfc(z / math.sqrt(2.0))
    return U1, p


def _label_sort_key(lbl):
    return lbl  # pandas groupby sorts labels as strings


def evaluate(lnL, labels, options_min_sim=25):
    """Statistics of sonics.pyx:136-181 on the cumulative results.

    lnL: float array; labels: array of 'a/b' strings (same length).
    Returns dict(min_sim, evaluated, best, second, best_ratio, pct_ratio, high_pval, n_tests)."""
    lnL = np.asarray(lnL, dtype=np.float64)
    labels = np.asarray(labels)
    groups = {}
    for g in sorted(set(labels.tolist()), key=_label_sort_key):
        groups[g] = lnL[labels == g]
    min_sim = min(int((v > SENTINEL).sum()) for v in groups.values())
    res = {"min_sim": min_sim, "evaluated": False, "groups": groups}
    if min_sim < options_min_sim:
        return res
    # sort groups by max lnL descending (stable over the label-sorted order)
    names = list(groups.keys())
    maxs = np.array([groups[g].max() for g in names])
    order = np.argsort(-maxs, kind="mergesort")
    best, second = names[order[0]], names[order[1]]
    best_ratio = maxs[order[0]] - maxs[order[1]]
    sb = np.sort(groups[best])
    ss = np.sort(groups[second])
    pct_ratio = quantile_linear(sb, 0.75) - quantile_linear(ss, 0.75)
    high = 0
    n_tests = 0
    for g in names:
        if g == best:
            continue
        try:
            _, p = mwu_greater(groups[best], groups[g])
            high = max(p, high)
        except ValueError:
            high = 1
        n_tests += 1
    high = high * n_tests
    res.update(evaluated=True, best=best, second=second, best_ratio=float(best_ratio),
               pct_ratio=float(pct_ratio), high_pval=high, n_tests=n_tests, sorted_best=sb, sorted_second=ss)
    return res


def _fmt(v):
    """str() as the reference applies it to numpy / Python scalars."""
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


def monte_carlo(max_n_reps, draw, padjust=0.001, lnL_threshold=2.3, min_sim_opt=25, monte_carlo_mode=False,
                add_ratios=""):
    """Restatement of monte_carlo() (sonics.pyx:47-269).

    draw(n) must return the NEXT n simulation records as a structured array with
    fields identity, r_squared, lnL, label (str) -- in the reference these come from
    one_repeat().  Returns (row_string, info_dict)."""
    recs = []
    run_reps = 0
    rnd = 1
    successful = False
    reps = max_n_reps if monte_carlo_mode else (100 if max_n_reps > 100 else max_n_reps)
    ev = None
    while run_reps < max_n_reps:
        recs.extend(draw(reps))
        run_reps += reps
        lnL = np.array([r["lnL"] for r in recs], dtype=np.float64)
        labels = np.array([r["label"] for r in recs])
        ident = np.array([r["identity"] for r in recs], dtype=np.float64)
        r2 = np.array([r["r_squared"] for r in recs], dtype=np.float64)
        if monte_carlo_mode:
            # :115-134  (sort_values descending, first row; ties resolved by first occurrence)
            bi = int(np.argmax(lnL))
            g = labels[bi]
            m = labels == g
            row = [g, ident[bi], quantile_linear(np.sort(ident[m]), 0.75), r2[bi],
                   quantile_linear(np.sort(r2[m]), 0.75), lnL[bi], quantile_linear(np.sort(lnL[m]), 0.75), run_reps]
            return "\t".join(_fmt(e) for e in row), {"genotype": g, "run_reps": run_reps}
        ev = evaluate(lnL, labels, min_sim_opt)
        if ev["evaluated"]:
            if ev["high_pval"] < padjust and ev["pct_ratio"] > lnL_threshold and ev["best_ratio"] > lnL_threshold:
                successful = True
                break
        reps = 4 * run_reps if rnd % 2 == 1 else run_reps
        if reps + run_reps > max_n_reps:
            reps = max_n_reps - run_reps
        rnd += 1

    if ev["min_sim"] < 25:  # hard-coded in the reference (:208)
        row = ["./.", ".", ".", ".", "no_success", ".", ".", ".", str(run_reps), "."]
        return "\t".join(row), {"genotype": "./.", "filter": "no_success", "run_reps": run_reps}
    if not ev["evaluated"]:
        # 25 <= min_sim < --min_sim: the reference dies with an unbound local
        raise NameError("high_pval referenced before assignment (reference behaviour for min_sim option > 25)")
    high = ev["high_pval"] if ev["high_pval"] < 1 else 1
    best = ev["best"]
    m = labels == best
    idx = np.nonzero(m)[0]
    bi = idx[int(np.argmax(lnL[idx]))]
    pct_ratio = ev["pct_ratio"]
    add_data = "."
    if add_ratios != "":
        for perc in add_ratios.split(";"):
            p = float(perc)
            pct_ratio = quantile_linear(ev["sorted_best"], p) - quantile_linear(ev["sorted_second"], p)
            if add_data == ".":
                add_data = "{}|{}".format(perc, _fmt(pct_ratio))
            else:
                add_data += ";{}|{}".format(perc, _fmt(pct_ratio))
    if successful:
        filt = "PASS"
    else:
        conds = ["MWU_test" if high > padjust else "",
                 "best_ratio" if ev["best_ratio"] < lnL_threshold else "",
                 "percentile_ratio" if pct_ratio < lnL_threshold else ""]
        filt = ",".join(c for c in conds if c != "")
    row = [best, ident[bi], r2[bi], lnL[bi], filt, high, ev["best_ratio"], pct_ratio, run_reps, add_data]
    info = {"genotype": best, "filter": filt, "run_reps": run_reps, "high_pval": high,
            "best_ratio": ev["best_ratio"], "pct_ratio": pct_ratio, "best_index": int(bi), "lnL": float(lnL[bi])}
    return "\t".join(_fmt(e) for e in row), info
