"""Consumer of the batch CSVs (reference viz/plot_runtime_boxplot.py:26-122): groups ``time_sec``
by ``N`` over successful runs and draws a log-y box plot.  Out of scope for acceleration."""

import csv
from collections import defaultdict


def read_times(csv_path):
    by_n = defaultdict(list)
    with open(csv_path, newline="", encoding="utf-8") as f:
        for row in csv.DictReader(f):
            if row.get("status") == "success":
                by_n[int(row["N"])].append(float(row["time_sec"]))
    return dict(sorted(by_n.items()))


def make_boxplot(csv_path, save_path=None, show=False):
    by_n = read_times(csv_path)
    try:
        import matplotlib.pyplot as plt
    except Exception:
        print("matplotlib not available: skipping plot")
        return by_n
    fig, ax = plt.subplots()
    ax.boxplot(list(by_n.values()), labels=[str(n) for n in by_n])
    ax.set_yscale("log")
    ax.set_xlabel("N")
    ax.set_ylabel("time_sec")
    if save_path:
        fig.savefig(save_path)
    if show:
        plt.show()
    return by_n


# same module-level configuration and zero-argument entry point as the reference (plot_runtime_boxplot.py:19-22, 118-120)
CONFIG = {
    "data_dir": "results/trial_2",          # folder with scp_benchmark_*.csv
    "out_path": "plots/scp_boxplot.pdf",    # where to save the plot
}


def main():
    import glob
    import os

    files = sorted(glob.glob(os.path.join(CONFIG["data_dir"], "scp_benchmark_*.csv")))
    if not files:
        raise FileNotFoundError(f"No 'scp_benchmark_*.csv' files in {CONFIG['data_dir']}")
    merged = {}
    for fp in files:
        for n, ts in read_times(fp).items():
            merged.setdefault(n, []).extend(ts)
    os.makedirs(os.path.dirname(CONFIG["out_path"]) or ".", exist_ok=True)
    make_boxplot(files[-1], save_path=CONFIG["out_path"])
    print(f"Saved plot: {CONFIG['out_path']}")
    return merged
