"""Host-side plotting pass-throughs (reference scp.py:644-840).  matplotlib is imported lazily;
when it is not installed the call is a no-op with a note (plots are not on the compute path)."""


def _plt():
    try:
        import matplotlib.pyplot as plt

        return plt
    except Exception:
        print("matplotlib not available: skipping plot")
        return None


def trajectories(solver, show_animation=False, save_path="trajectories.pdf"):
    plt = _plt()
    if plt is None:
        return None
    pos = solver.trajectories["positions"]
    fig, ax = plt.subplots(figsize=(7, 7))
    for i in range(pos.shape[0]):
        ax.plot(pos[i, :, 0], pos[i, :, 1], lw=1.2)
        ax.plot(pos[i, 0, 0], pos[i, 0, 1], "o", ms=4)
        ax.plot(pos[i, -1, 0], pos[i, -1, 1], "x", ms=5)
    ax.set_xlim(solver.space_dims[0], solver.space_dims[2])
    ax.set_ylim(solver.space_dims[1], solver.space_dims[3])
    ax.set_aspect("equal")
    if save_path:
        fig.savefig(save_path)
    if show_animation:
        plt.show()
    return fig


def snapshots(solver, num_snapshots=5, save_path=None):
    plt = _plt()
    if plt is None:
        return None
    pos = solver.trajectories["positions"]
    K = pos.shape[1]
    fig, axes = plt.subplots(1, num_snapshots, figsize=(4 * num_snapshots, 4))
    for a, k in zip(axes, [int(round(t * (K - 1) / max(1, num_snapshots - 1))) for t in range(num_snapshots)]):
        a.scatter(pos[:, k, 0], pos[:, k, 1], s=12)
        a.set_title(f"t = {k * solver.h:.1f} s")
        a.set_xlim(solver.space_dims[0], solver.space_dims[2])
        a.set_ylim(solver.space_dims[1], solver.space_dims[3])
        a.set_aspect("equal")
    if save_path:
        fig.savefig(save_path)
    return fig
